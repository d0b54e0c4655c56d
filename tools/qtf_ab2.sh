#!/bin/bash
# tile variant of the second-order force kernel vs the diagonal kernel: parity tests under both, then timing of both
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tiles"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "second_order or slender" 2>&1 | tail -6
echo "== diag"; RAFTK_QTF_DIAG=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "second_order" 2>&1 | tail -3
echo "== timing tiles"; timeout 300 python tools/qtf_timing.py 2>&1 | tail -1 | tee gpurun_out/qtf_timing_tiles.json
echo "== timing diag"; RAFTK_QTF_DIAG=1 timeout 300 python tools/qtf_timing.py 2>&1 | tail -1 | tee gpurun_out/qtf_timing_diag.json
