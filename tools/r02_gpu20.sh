#!/bin/bash
# k_farm_rows with one division per elimination step and reciprocal diagonals in the back substitution: tests + timing
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_farm.py -q -m gpu 2>&1 | tail -2
timeout 300 python __graft_entry__.py smoke 2>&1 | sed -n 2p
timeout 600 python bench.py --workload farm --steps 10 --warmup 3 --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']; p = d.get('parity') or {}
print('farm N=2: ms/step %.4f value %.4g e2e %.4g system kernel %.4f ms parity %s/%s' % (d['ms_per_step'], d['value'], d['e2e']['value'], r['kernel_ms'], p.get('max_rel_err'), p.get('pass_mismatch_units')))"
