#!/bin/bash
# last GPU pass of the round: full GPU suite + smoke on the final tree
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu21.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/r02_smoke21.txt
