#!/bin/bash
# A/B: threads per CTA of the blocked shared-memory LU (farm, 6N = 48 and 96)
cd "$(dirname "$0")/.."
for n in 8 16; do for t in 256 512 1024; do
  RAFTK_FARM_THREADS=$t timeout 600 python bench.py --workload farm --turbines $n --cases 16 --steps 3 --warmup 3 --no-extras --no-parity 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
print('farm N=$n threads $t: system kernel %.3f ms lu %.0f GF/s step %.3f ms' % (r['kernel_ms'], r['lu_gflops'], d['ms_per_step']))"
done; done
