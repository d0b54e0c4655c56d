#!/bin/bash
# N=2 diagnostic: why does bench.py's step (0.436 ms) exceed solve + barrier as timed by tools/peer_probe.py (0.324 ms)?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 \
      bench.py --gpus 2 --steps $2 --warmup $3 --no-extras --no-e2e --no-parity $4 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('$5 steps $2 warmup $3: ms %.4f' % l['ms_per_step'], l['per_rank']['solve_ms'], 'skew %.4f' % l['per_rank']['exchange_and_skew_ms'], l['clocks']['samples'])"
}
run 29521 20 5 "" default
RAFTK_BENCH_NO_SAMPLER=1 run 29522 20 5 "" nosampler
run 29523 20 40 "" warm40
run 29524 200 5 "" steps200
run 29525 20 5 "--exchange nccl" nccl
