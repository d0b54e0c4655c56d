#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/slender_timing.py 2>&1 | tail -1 | tee gpurun_out/slender_timing.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/slender_launches.csv python tools/slender_timing.py > /dev/null 2>&1
timeout 300 python bench.py --workload cfg3q --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cfg3q.json
timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cfg3.json
