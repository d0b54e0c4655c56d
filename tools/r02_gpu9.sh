#!/bin/bash
# round 2, ninth GPU pass (re-entry check at HEAD): full GPU suite + smoke + default bench (with CPU baselines) + reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02_pytest_gpu9.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench9.json 2> gpurun_out/r02_bench9.err
tail -c 1500 gpurun_out/r02_bench9.json | head -c 600; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench9_ref.json 2>> gpurun_out/r02_bench9.err
tail -c 800 gpurun_out/r02_bench9_ref.json; tail -3 gpurun_out/r02_bench9.err
