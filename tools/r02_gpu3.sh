#!/bin/bash
# round 2, third GPU pass: validate k_rao_fused2 (sanitizer on a small case, full GPU suite), A/B against the first-generation kernel, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/fused2_check.py 256 3 0 2>&1 | tail -3
timeout 300 python tools/fused2_check.py 512 4 2 2>&1 | tail -3
timeout 300 python tools/fused2_check.py 1000 5 4 2>&1 | tail -3
timeout 600 compute-sanitizer --tool memcheck python tools/fused2_check.py 256 2 0 2>&1 | tail -8 | tee gpurun_out/r02_sanitizer_fused2.txt
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu.txt
for v in gen1 gen2; do
  if [ $v = gen1 ]; then export RAFTK_FUSED_GEN1=1; else unset RAFTK_FUSED_GEN1; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02_ab_$v.json 2> gpurun_out/r02_ab_$v.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/r02_ab_$v.json").read().strip().split("\n")[-1])
print("$v", "ms/step %.4f" % l["ms_per_step"], "value %.4g" % l["value"], "e2e %.4g" % l["e2e"]["value"], "kernel_ms %.4f" % l["roofline"]["kernel_ms"], "fp64 frac", l["roofline_fp64"]["frac"], "parity", l["parity"]["max_rel_err"], l["parity"]["pass_mismatch_units"], "launches", l["gpu_launches"])
PY
  tail -2 gpurun_out/r02_ab_$v.err
done
unset RAFTK_FUSED_GEN1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rao_fused2 -c 1 -s 3 -o gpurun_out/r02_fused2_a -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-extras > gpurun_out/r02_ncu_fused2.log 2>&1
tail -2 gpurun_out/r02_ncu_fused2.log | cut -c1-300
