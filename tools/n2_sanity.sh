#!/bin/bash
# N = 2 sanity of the driver's launch line (default workload) + the reference arm under torchrun
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_n2.json | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
