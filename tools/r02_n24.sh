#!/bin/bash
# 8-GPU box, final kernel: N = 4 and N = 2 (fused exchange), driver's torchrun line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n \
    bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_scale_s2_n$n.json 2> gpurun_out/r02_scale_s2_n$n.err
echo "== N=$n rc=$?"
done
python - <<PY
import json
for n in (4, 2):
    try:
        l=json.loads(open("gpurun_out/r02_scale_s2_n%d.json" % n).read().strip().split("\n")[-1])
        s=l.get("sweep") or {}
        print("N=%d: cfg2 ms %.4f value %.4g e2e_ms %.4f e2e %.4g | sweep ms %.3f value %.4g e2e %.4g incl_build %.4g | verified %s" % (n, l["ms_per_step"], l["value"], l["e2e"]["ms_per_step"], l["e2e"]["value"], s.get("ms_per_step", 0), s.get("value", 0), (s.get("e2e") or {}).get("value", 0), (s.get("e2e_including_table_build") or {}).get("value", 0), l.get("exchange_verified")))
        print("   per_rank", json.dumps(l.get("per_rank")))
    except Exception as e:
        print("failed", n, e)
PY
