#!/bin/bash
# GPU check of the second-order-force row: its parity tests, a timing line, and the unchanged default bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "second_order" 2>&1 | tail -15
timeout 300 python tools/qtf_timing.py 2>&1 | tail -3 | tee gpurun_out/qtf_timing.json
timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_after_qtf.json
