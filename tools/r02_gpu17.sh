#!/bin/bash
# N=2 check of the batched host->device staging in ShardedSolve.step_host + struct caching (e2e at N > 1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_exchange.py -q -m gpu 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/r02_n2d.json 2> gpurun_out/r02_n2d.err
echo "rc=$?"; python - <<PY
import json
l=json.loads(open("gpurun_out/r02_n2d.json").read().strip().split("\n")[-1])
print("N=2 ms %.4f value %.4g e2e_ms %.4f e2e %.4g verified %s warm %s" % (l["ms_per_step"], l["value"], l["e2e"]["ms_per_step"], l["e2e"]["value"], l.get("exchange_verified"), l["config"].get("warmup_steps_run")))
print(json.dumps(l.get("per_rank")))
PY
tail -2 gpurun_out/r02_n2d.err | cut -c1-300
