#!/bin/bash
# round 2, final 1-GPU validation of the second session: full GPU suite, smoke, default bench (all keys), reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02_pytest_gpu18.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/r02_smoke18.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench18.json 2> gpurun_out/r02_bench18.err
python - <<PY
import json
l=json.loads(open("gpurun_out/r02_bench18.json").read().strip().split("\n")[-1])
print("cfg2 ms/step %.4f value %.4g e2e %.4g (%.4f ms) fp64 %s parity %.2e/%d launches %d sustained %.4g" % (l["ms_per_step"], l["value"], l["e2e"]["value"], l["e2e"]["ms_per_step"], l["roofline_fp64"]["frac"], l["parity"]["max_rel_err"], l["parity"]["pass_mismatch_units"], l["gpu_launches"], l["sustained"]["value"]))
s=l["sweep"]; print("sweep ms/step %.3f value %.4g e2e %.4g incl build %.4g (build %.3f s) fp64 %s parity %.2e/%d" % (s["ms_per_step"], s["value"], s["e2e"]["value"], s["e2e_including_table_build"]["value"], s["config"]["table_build_s"], s["roofline_fp64"]["frac"], s["parity"]["max_rel_err"], s["parity"]["pass_mismatch_units"]))
print("cpu", l["cpu_baseline"]["value"], l["cpu_baseline"].get("reference_numpy", {}).get("single_process"))
PY
tail -3 gpurun_out/r02_bench18.err
