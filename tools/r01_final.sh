#!/bin/bash
# round-end validation: full GPU parity suite, smoke, default bench + reference arm, cfg3q with the tile kernel, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.txt
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench_cfg2.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; tail -c 400 gpurun_out/bench_reference.json
timeout 300 python bench.py --workload cfg3q --steps 5 --warmup 3 --no-cpu-baseline 2>> gpurun_out/bench.err | tail -1 > gpurun_out/bench_cfg3q.json; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3q.json')); print('cfg3q ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -2 gpurun_out/bench.err
