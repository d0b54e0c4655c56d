#!/bin/bash
# round 2, twelfth GPU pass: A/B of the k_rao_fused2 switches (one per build) on cfg2 and the sweep shard
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ab() {
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extras $2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
p = d.get('parity') or {}
print('$1 $2', 'ms/step %.4f' % d['ms_per_step'], 'value %.4g' % d['value'], 'fp64 %s' % d['roofline_fp64']['frac'], 'parity %s/%s' % (p.get('max_rel_err'), p.get('pass_mismatch_units')))"
}
for lib in "$@"; do
  export RAFTK_LIB="$PWD/build_ab/$lib.so"
  ab $lib ""
  ab $lib "--workload sweep --steps 3"
done 2>&1 | tee gpurun_out/r02_ab12.txt
