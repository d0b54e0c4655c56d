#!/bin/bash
# A/B of k_qtf_force builds (register cap / block size) + the second-order parity tests on the in-tree build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "second_order" 2>&1 | tail -8
for lib in default build/ab/q1.so build/ab/q2.so build/ab/q3.so build/ab/q4.so; do
  if [ "$lib" = default ]; then unset RAFTK_LIB; else export RAFTK_LIB="$PWD/$lib"; fi
  echo "== $lib"; timeout 300 python tools/qtf_timing.py 2>&1 | tail -1 | tee -a gpurun_out/qtf_ab.txt
done
