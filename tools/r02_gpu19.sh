#!/bin/bash
# ncu --set full capture of k_farm_rows<12> (64 sea states x 1024 bins, N = 2)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_farm_rows -s 3 -c 1 -o gpurun_out/r02_farm_rows -f \
  python bench.py --workload farm --steps 2 --warmup 3 --no-extras --no-parity > /dev/null 2>&1
ls -la gpurun_out/r02_farm_rows.ncu-rep
