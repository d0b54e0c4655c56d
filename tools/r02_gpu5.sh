#!/bin/bash
# round 2, fifth GPU pass: new farm kernels (warp-per-system / blocked) and the lighter generalised-DOF LU: tests, sanitizer, bench lines, smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu.txt
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_farm.py tests/test_general_dofs.py -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02_sanitizer_new.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/r02_smoke.txt
for wl in farm flex; do
  timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_$wl.json 2> gpurun_out/r02_bench_$wl.err
  python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/r02_bench_$wl.json").read().strip().split("\n")[-1])
    print("$wl", "ms/step %.4f value %.4g e2e %.4g" % (l["ms_per_step"], l["value"], l["e2e"]["value"]), "parity", (l.get("parity") or {}).get("max_rel_err"), (l.get("parity") or {}).get("pass_mismatch_units"), "kernel_ms", l["roofline"].get("kernel_ms"), (l.get("roofline_fp64") or {}).get("frac"), json.dumps(l.get("farm_sizes")))
except Exception as e:
    print("$wl failed", e)
PY
  tail -3 gpurun_out/r02_bench_$wl.err
done
