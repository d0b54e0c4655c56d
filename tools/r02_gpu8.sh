#!/bin/bash
# round 2, eighth GPU pass: projections kept across chunks + pairwise transcendentals: full GPU suite + default bench + cfg3 + sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02_pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err
python - <<PY
import json
l=json.loads(open("gpurun_out/r02_bench_d.json").read().strip().split("\n")[-1])
print("cfg2 ms/step %.4f value %.4g e2e %.4g kernel_ms %.4f fp64 %s parity %.2e/%d launches %d sustained %.4g" % (l["ms_per_step"], l["value"], l["e2e"]["value"], l["roofline"]["kernel_ms"], l["roofline_fp64"]["frac"], l["parity"]["max_rel_err"], l["parity"]["pass_mismatch_units"], l["gpu_launches"], l["sustained"]["value"]))
s=l["sweep"]; print("sweep ms/step %.3f value %.4g e2e %.4g parity %.2e/%d" % (s["ms_per_step"], s["value"], s["e2e"]["value"], s["parity"]["max_rel_err"], s["parity"]["pass_mismatch_units"]))
PY
tail -2 gpurun_out/r02_bench_d.err
RAFTK_FUSED_GEN1=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-e2e --no-parity 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('gen1 on this box: ms/step %.4f' % l['ms_per_step'])"
timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('cfg3 ms/step %.4f value %.4g e2e %.4g parity %s' % (l['ms_per_step'], l['value'], l['e2e']['value'], l['parity']['max_rel_err']))"
