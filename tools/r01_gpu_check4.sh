#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "slender or second_order" 2>&1 | tail -5
timeout 300 python tools/slender_timing.py 2>&1 | tail -1 | tee gpurun_out/slender_timing.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/slender_launches.csv python tools/slender_timing.py > /dev/null 2>&1
