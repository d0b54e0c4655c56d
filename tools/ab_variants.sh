#!/bin/bash
# A/B timing of alternative builds of libraftk.so (same ABI) on the default bench workload.
# usage: tools/ab_variants.sh <cluster> <steps> lib1.so lib2.so ...   ("default" = in-tree library)
cs=$1; steps=$2; shift 2
for lib in "$@"; do
  if [ "$lib" = default ]; then unset RAFTK_LIB; else export RAFTK_LIB="$PWD/$lib"; fi
  python bench.py --steps "$steps" --warmup 5 --no-cpu-baseline --no-e2e --cluster "$cs" 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'])"
done
