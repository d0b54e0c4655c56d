#!/bin/bash
# round 2, first GPU pass: full GPU suite, smoke, default bench, reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/r02_smi.txt
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/r02_pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/r02_smoke.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -c 3000 gpurun_out/r02_bench_n1.json; tail -5 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
tail -c 2000 gpurun_out/r02_bench_ref.json; tail -5 gpurun_out/r02_bench_ref.err
