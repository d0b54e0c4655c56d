#!/bin/bash
# smallest possible GPU check of the sweep workload built by the native table builder
cd "$(dirname "$0")/.."
timeout 60 python bench.py --workload sweep --designs 96 --steps 2 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); p = d['parity']
print('sweep(96) ms/step %.3f builder %s build %.4f/%.4f s parity %s/%s ok %s' % (d['ms_per_step'], d['config']['table_builder'][:28], d['config']['table_build_s'], d['config']['table_build_warm_s'], p['max_rel_err'], p['pass_mismatch_units'], p['ok']))"
