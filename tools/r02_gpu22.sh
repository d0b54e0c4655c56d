#!/bin/bash
# sanity of the sweep workload after the builder changes (tables from batch_builder, parity vs the oracle on per-design tables)
cd "$(dirname "$0")/.."
timeout 300 python bench.py --workload sweep --steps 2 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); p = d['parity']
print('sweep ms/step %.3f value %.4g e2e %.4g build %.3f s parity %s/%s ok %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['table_build_s'], p['max_rel_err'], p['pass_mismatch_units'], p['ok']))"
