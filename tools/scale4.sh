#!/bin/bash
N=${1:-4}
for wl in "cfg2 --steps 30" "sweep --designs 600 --steps 3 --no-e2e"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --workload $wl --warmup 3 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip())
print('N=$N', d['config']['workload'][:18], 'value %.4e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], (d.get('e2e') or {}).get('value'))"
done
python bench.py --workload sweep --designs 600 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip())
print('N=1 sweep600', 'value %.4e' % d['value'], 'ms/step %.4f' % d['ms_per_step'])"
