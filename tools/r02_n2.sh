#!/bin/bash
# N=2: fused exchange (CUDA IPC peer stores) vs NCCL all-gather, driver's torchrun line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for ex in fused nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 20 --warmup 5 --exchange $ex > gpurun_out/r02_bench_n2_$ex.json 2> gpurun_out/r02_bench_n2_$ex.err
  echo "== $ex rc=$?"; tail -c 1500 gpurun_out/r02_bench_n2_$ex.json; tail -3 gpurun_out/r02_bench_n2_$ex.err
done
