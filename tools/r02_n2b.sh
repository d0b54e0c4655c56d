#!/bin/bash
# N=2 sanity of the final kernel: driver's torchrun line (fused exchange), per-rank block
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/r02_n2b.json 2> gpurun_out/r02_n2b.err
echo "rc=$?"; python - <<PY
import json
l=json.loads(open("gpurun_out/r02_n2b.json").read().strip().split("\n")[-1])
print("N=2 ms %.4f value %.4g e2e_ms %.4f verified %s" % (l["ms_per_step"], l["value"], l["e2e"]["ms_per_step"], l.get("exchange_verified")))
print(json.dumps(l.get("per_rank")))
PY
tail -3 gpurun_out/r02_n2b.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/peer_probe.py 2>&1 | tail -12
