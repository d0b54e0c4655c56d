#!/bin/bash
# round 2, fifteenth GPU pass: fixed masks test + farm tests, farm bench record (all sizes), memcheck over smoke() (fused2 with
# direction masks, k_farm_rows), racecheck of the two-bin solver's shared-memory exchanges
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused2.py tests/test_farm.py -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --workload farm --steps 10 --warmup 3 > gpurun_out/r02_bench_farm_s2.json 2> gpurun_out/r02_bench_farm_s2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_farm_s2.json").read().strip().split("\n")[-1])
r=d["roofline"]; p=d.get("parity") or {}
print("farm N=2: ms/step %.4f value %.4g e2e %.4g system kernel %.4f ms parity %s/%s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], r["kernel_ms"], p.get("max_rel_err"), p.get("pass_mismatch_units")))
print("file case", d["file_case"]["value"], d["file_case"]["ms_per_step"])
for k, v in d["farm_sizes"].items(): print(" N =", k, {a: round(b, 4) if b < 1e4 else float("%.4g" % b) for a, b in v.items()})
PY
tail -2 gpurun_out/r02_bench_farm_s2.err
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py smoke > gpurun_out/r02_sanitizer_s2.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_sanitizer_s2.txt
tail -6 gpurun_out/r02_sanitizer_s2.txt
