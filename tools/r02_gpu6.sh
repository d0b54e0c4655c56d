#!/bin/bash
# round 2, sixth GPU pass: LU A/B micro-benchmark, ncu summaries (cfg3 fused2, farm, general LU v2, slender pairs), launch list of the default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
./tools/micro/lu6_ab.bin 2>&1 | tee gpurun_out/r02_lu6_ab.txt
timeout 600 python -m pytest tests/test_farm.py tests/test_exchange.py -x -q -m gpu 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_default.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02_launch_run.log 2>&1
grep -c "k_rao_fused2\|k_fused_plan" gpurun_out/r02_launches_default.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rao_fused2 -c 1 -s 3 -o gpurun_out/r02_fused2_cfg3 -f \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-extras > gpurun_out/r02_ncu_cfg3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_farm_response" -c 3 -o gpurun_out/r02_farm -f \
    python bench.py --workload farm --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-extras > gpurun_out/r02_ncu_farm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gen_solve_blocked" -c 1 -s 2 -o gpurun_out/r02_flex_lu2 -f \
    python bench.py --workload flex --steps 1 --warmup 1 --no-parity --cases 16 > gpurun_out/r02_ncu_flex2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_slender_pairs" -c 1 -o gpurun_out/r02_slender -f \
    python tools/slender_timing.py > gpurun_out/r02_ncu_slender.log 2>&1
for wl in farm; do
  timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_$wl.json 2> gpurun_out/r02_bench_$wl.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r02_bench_$wl.json").read().strip().split("\n")[-1])
print("$wl", "ms/step %.4f value %.4g" % (l["ms_per_step"], l["value"]), json.dumps(l.get("farm_sizes")))
PY
done
ls -la gpurun_out/*.ncu-rep | tail -5
