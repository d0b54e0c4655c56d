#!/bin/bash
# usage: tools/scale_check.sh N  -- compare pipelined (2 chunks) vs single all-gather at N GPUs, both workloads
N=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:3}" 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$2', 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'])"; }
run 29531 "N$N cfg2 chunks2" --steps 30 --warmup 3 --no-e2e
run 29532 "N$N cfg2 chunks1" --steps 30 --warmup 3 --no-e2e --chunks 1
run 29533 "N$N sweep chunks2" --workload sweep --designs 300 --steps 5 --warmup 3 --no-e2e --chunks 2
run 29534 "N$N sweep chunks4" --workload sweep --designs 300 --steps 5 --warmup 3 --no-e2e --chunks 4
run 29535 "N$N sweep chunks1" --workload sweep --designs 300 --steps 5 --warmup 3 --no-e2e --chunks 1
python bench.py --workload sweep --designs 300 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N1 sweep', 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'])"
