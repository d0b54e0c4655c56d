#!/bin/bash
# round 2, fourteenth GPU pass: register-row farm kernel (6N = 12, 18, 24) vs the shared-memory warp kernel; new tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_farm.py tests/test_fused2.py -x -q -m gpu 2>&1 | tail -5
fb() {
  timeout 600 python bench.py --workload farm --turbines $1 --steps 10 --warmup 3 --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']; p = d.get('parity') or {}
print('farm N=$1 $2: ms/step %.4f value %.4g e2e %.4g (%.3f ms) system kernel %.4f ms lu %.0f GF/s parity %s/%s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], r['kernel_ms'], r['lu_gflops'], p.get('max_rel_err'), p.get('pass_mismatch_units')))"
}
for n in 2 4; do
  fb $n rows
  RAFTK_FARM_SMEM=1 fb $n smem
done
fb 3 rows
