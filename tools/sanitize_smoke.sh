#!/bin/bash
# compute-sanitizer memcheck over smoke(): fused solve + second-order force (tile kernel) + slender-body QTF + channel stats
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py smoke > gpurun_out/sanitizer_smoke.txt 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_smoke.txt
tail -8 gpurun_out/sanitizer_smoke.txt
