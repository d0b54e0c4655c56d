#!/bin/bash
# round 2, eleventh GPU pass: direction masks + merged barriers (one cluster barrier per exchange) + tree sums + sqrt-free
# convergence test: parity suite, A/B against the previous build and the part-2 unroll variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02_pytest_gpu11.txt
ab() {
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extras $2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
p = d.get('parity') or {}
print('$1 $2', 'ms/step %.4f' % d['ms_per_step'], 'value %.4g' % d['value'], 'fp64 %s' % d['roofline_fp64']['frac'], 'parity %s/%s' % (p.get('max_rel_err'), p.get('pass_mismatch_units')))"
}
for lib in build_ab/base.so default build_ab/p2u1.so build_ab/p2u4.so; do
  if [ "$lib" = default ]; then unset RAFTK_LIB; else export RAFTK_LIB="$PWD/$lib"; fi
  ab $lib ""
  ab $lib "--workload sweep --steps 3"
done
unset RAFTK_LIB
ab default "--workload cfg3 --steps 10"
