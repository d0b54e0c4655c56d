#!/bin/bash
# Full GPU parity suite + smoke + ncu evidence for the second-order force kernel (launch list, one --set full capture).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/qtf_launches.csv python tools/qtf_timing.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_qtf_force -s 2 -c 1 -o gpurun_out/prof_qtf python tools/qtf_timing.py > /dev/null 2>&1
ls -la gpurun_out/ | tail -8
