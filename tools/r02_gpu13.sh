#!/bin/bash
# round 2, thirteenth GPU pass: final k_rao_fused2 (direction masks, parallel remote reads, one flag barrier):
# full GPU suite + smoke, default bench (all keys), cfg3, reference arm, ncu launch list + full capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02_pytest_gpu13.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/r02_smoke13.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench13.json 2> gpurun_out/r02_bench13.err
timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench13_cfg3.json 2>> gpurun_out/r02_bench13.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench13_ref.json 2>> gpurun_out/r02_bench13.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches13.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rao_fused2 -c 1 -s 3 -o gpurun_out/r02_fused2_final -f \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-extras > /dev/null 2>&1
python - <<PY
import json
l=json.loads(open("gpurun_out/r02_bench13.json").read().strip().split("\n")[-1])
print("cfg2 ms/step %.4f value %.4g e2e %.4g (%.4f ms) fp64 %s parity %.2e/%d launches %d sustained %.4g" % (l["ms_per_step"], l["value"], l["e2e"]["value"], l["e2e"]["ms_per_step"], l["roofline_fp64"]["frac"], l["parity"]["max_rel_err"], l["parity"]["pass_mismatch_units"], l["gpu_launches"], l["sustained"]["value"]))
s=l["sweep"]; print("sweep ms/step %.3f value %.4g e2e %.4g fp64 %s parity %.2e/%d" % (s["ms_per_step"], s["value"], s["e2e"]["value"], s["roofline_fp64"]["frac"], s["parity"]["max_rel_err"], s["parity"]["pass_mismatch_units"]))
c=json.loads(open("gpurun_out/r02_bench13_cfg3.json").read().strip().split("\n")[-1])
print("cfg3 ms/step %.4f value %.4g e2e %.4g fp64 %s parity %s" % (c["ms_per_step"], c["value"], c["e2e"]["value"], c["roofline_fp64"]["frac"], c["parity"]["max_rel_err"]))
PY
tail -3 gpurun_out/r02_bench13.err
