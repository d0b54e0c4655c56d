#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/peer_probe.py 2>&1 | grep -v "^\*\|OMP_NUM" | tee gpurun_out/r02_peer_probe_n2.txt
