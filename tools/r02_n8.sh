#!/bin/bash
# 8-GPU box: the driver's torchrun line at N = 8 (fused exchange, then NCCL all-gather for the A/B), N = 4 and N = 2 (fused)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # n exchange tag
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2952$1 \
      bench.py --gpus $1 --steps 20 --warmup 5 --exchange $2 > gpurun_out/r02_scale_n$1_$2.json 2> gpurun_out/r02_scale_n$1_$2.err
  echo "== N=$1 $2 rc=$?"
  python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/r02_scale_n$1_$2.json").read().strip().split("\n")[-1])
    s=l.get("sweep") or {}
    print("N=$1 $2: cfg2 ms %.4f value %.4g e2e_ms %.4f e2e %.4g kernel %.4f sustained_ms %.4f | sweep ms %.3f value %.4g e2e %.4g incl_build %.4g | verified %s" % (l["ms_per_step"], l["value"], l["e2e"]["ms_per_step"], l["e2e"]["value"], l["roofline"]["kernel_ms"], l["sustained"]["ms_per_step"], s.get("ms_per_step", 0), s.get("value", 0), (s.get("e2e") or {}).get("value", 0), (s.get("e2e_including_table_build") or {}).get("value", 0), l.get("exchange_verified")))
except Exception as e:
    print("failed", e)
PY
  tail -2 gpurun_out/r02_scale_n$1_$2.err | cut -c1-300
}
run 8 fused
run 8 nccl
run 4 fused
run 2 fused
timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err
python - <<PY
import json
l=json.loads(open("gpurun_out/r02_scale_n1.json").read().strip().split("\n")[-1]); s=l["sweep"]
print("N=1: cfg2 ms %.4f value %.4g e2e %.4g | sweep ms %.3f value %.4g" % (l["ms_per_step"], l["value"], l["e2e"]["value"], s["ms_per_step"], s["value"]))
PY
