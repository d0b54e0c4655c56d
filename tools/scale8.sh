#!/bin/bash
# 8-GPU record: default workload (cfg2 weak scaling) and the full cfg4 sweep (10000 designs x 16 cases x 512 bins)
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 30 --warmup 3 2>/dev/null | tail -1 > gpurun_out/scale${N}_cfg2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --workload sweep --steps 3 --warmup 3 --no-e2e 2>/dev/null | tail -1 > gpurun_out/scale${N}_sweep.json
for f in cfg2 sweep; do python - <<PY
import json
d = json.loads(open('gpurun_out/scale${N}_$f.json').read().strip())
print('N=$N $f', 'value %.4e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], d['config']['collective'][:60], d['clocks'])
PY
done
