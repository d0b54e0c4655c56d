#!/bin/bash
# round 2, fourth GPU pass: full suite with farm + blocked general LU, sanitizer on the new kernels, bench default/farm/flex, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu.txt
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_farm.py tests/test_general_dofs.py tests/test_exchange.py -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02_sanitizer_new.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err
python - <<PY
import json
l=json.loads(open("gpurun_out/r02_bench_b.json").read().strip().split("\n")[-1])
print("cfg2 ms/step %.4f value %.4g e2e %.4g kernel_ms %.4f fp64 %s parity %.2e/%d launches %d sustained %.4g" % (l["ms_per_step"], l["value"], l["e2e"]["value"], l["roofline"]["kernel_ms"], l["roofline_fp64"]["frac"], l["parity"]["max_rel_err"], l["parity"]["pass_mismatch_units"], l["gpu_launches"], l["sustained"]["value"]))
s=l["sweep"]; print("sweep ms/step %.3f value %.4g e2e %.4g incl.build %.4g parity %.2e/%d" % (s["ms_per_step"], s["value"], s["e2e"]["value"], s["e2e_including_table_build"]["value"], s["parity"]["max_rel_err"], s["parity"]["pass_mismatch_units"]))
PY
tail -2 gpurun_out/r02_bench_b.err
for wl in cfg3 farm flex; do
  timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_$wl.json 2> gpurun_out/r02_bench_$wl.err
  python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/r02_bench_$wl.json").read().strip().split("\n")[-1])
    print("$wl", "ms/step %.4f value %.4g e2e %.4g" % (l["ms_per_step"], l["value"], l["e2e"]["value"]), "parity", (l.get("parity") or {}).get("max_rel_err"), (l.get("parity") or {}).get("pass_mismatch_units"), "roofline", l["roofline"].get("kernel_ms"), (l.get("roofline_fp64") or {}).get("frac"), l.get("farm_sizes"))
except Exception as e:
    print("$wl failed", e)
PY
  tail -3 gpurun_out/r02_bench_$wl.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rao_fused2 -c 1 -s 3 -o gpurun_out/r02_fused2_b -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-extras > gpurun_out/r02_ncu_fused2b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gen_solve_blocked|k_farm_response" -c 2 -s 2 -o gpurun_out/r02_flex_lu -f \
    python bench.py --workload flex --steps 1 --warmup 1 --no-parity --cases 16 > gpurun_out/r02_ncu_flex.log 2>&1
tail -1 gpurun_out/r02_ncu_flex.log | cut -c1-200
