#!/bin/bash
# full GPU parity suite + smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
