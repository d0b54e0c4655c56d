#!/bin/bash
# round 2, second GPU pass: full GPU suite, FP64 pipe micro-benchmarks, ncu capture of the fused solver (cfg2)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/r02_pytest_gpu.txt
./tools/micro/fp64_micro.bin 2>&1 | tee gpurun_out/r02_fp64_micro.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rao_fused -c 1 -s 3 -o gpurun_out/r02_fused_base -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-extras > gpurun_out/r02_ncu_base.log 2>&1
tail -3 gpurun_out/r02_ncu_base.log
ls -la gpurun_out/*.ncu-rep | tail -3
