#!/bin/bash
# round-end measurement set (1 GPU): default bench line, cfg3, sweep shard, ncu launch list + full capture
mkdir -p gpurun_out
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2>> gpurun_out/bench_cfg2.err
python bench.py --workload sweep --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sweep.json 2>> gpurun_out/bench_cfg2.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_cfg2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rao_fused -s 3 -c 1 -o gpurun_out/prof_fused_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
for f in cfg2 cfg3 sweep reference; do python - <<PY
import json
d = json.loads(open('gpurun_out/bench_$f.json').read().strip().splitlines()[-1])
e = d.get('e2e') or {}
print('$f', 'value %.4e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.4e' % e.get('value', float('nan')), d.get('clocks'))
PY
done
tail -3 gpurun_out/bench_cfg2.err
