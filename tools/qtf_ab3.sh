#!/bin/bash
# tile kernel tuning variants + one ncu --set full capture of the default build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in default build/ab/t1.so build/ab/t2.so build/ab/t3.so build/ab/t4.so; do
  if [ "$lib" = default ]; then unset RAFTK_LIB; else export RAFTK_LIB="$PWD/$lib"; fi
  echo "== $lib"; timeout 300 python tools/qtf_timing.py 2>&1 | tail -1 | tee -a gpurun_out/qtf_ab3.txt
done
unset RAFTK_LIB
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_qtf_tiles -s 2 -c 1 -o gpurun_out/prof_qtf_tiles python tools/qtf_timing.py > /dev/null 2>&1
ls -la gpurun_out | tail -4
