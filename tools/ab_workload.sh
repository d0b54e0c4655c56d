#!/bin/bash
# usage: tools/ab_workload.sh "<bench args>" cs1 cs2 ...   -- same workload at several cluster sizes
wl="$1"; shift
for cs in "$@"; do
  python bench.py $wl --warmup 3 --no-cpu-baseline --no-e2e --cluster "$cs" 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip())
print('cluster $cs', 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'])"
done
