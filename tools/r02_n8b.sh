#!/bin/bash
# 8-GPU box, final kernel: the driver's torchrun line at N = 8 (fused exchange) and the N = 1 line on the same box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 \
    bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_scale_s2_n8.json 2> gpurun_out/r02_scale_s2_n8.err
echo "== N=8 rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_scale_s2_n1.json 2> gpurun_out/r02_scale_s2_n1.err
echo "== N=1 rc=$?"
python - <<PY
import json
for n in (8, 1):
    try:
        l=json.loads(open("gpurun_out/r02_scale_s2_n%d.json" % n).read().strip().split("\n")[-1])
        s=l.get("sweep") or {}
        print("N=%d: cfg2 ms %.4f value %.4g e2e_ms %.4f e2e %.4g | sweep ms %.3f value %.4g e2e %.4g incl_build %.4g | verified %s" % (n, l["ms_per_step"], l["value"], l["e2e"]["ms_per_step"], l["e2e"]["value"], s.get("ms_per_step", 0), s.get("value", 0), (s.get("e2e") or {}).get("value", 0), (s.get("e2e_including_table_build") or {}).get("value", 0), l.get("exchange_verified")))
        print("   per_rank", json.dumps(l.get("per_rank")))
        print("   sweep per_rank", json.dumps(s.get("per_rank")))
    except Exception as e:
        print("failed", n, e)
PY
tail -2 gpurun_out/r02_scale_s2_n8.err | cut -c1-300
