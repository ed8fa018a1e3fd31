#!/usr/bin/env bash
# round 2, call 6: LDS.128 few-address microbenchmark + full GPU suite (mesh colliders on the device)
set -uo pipefail
mkdir -p gpurun_out
./scripts/micro/lds_groups > gpurun_out/r02h_lds_groups.txt 2>&1
cat gpurun_out/r02h_lds_groups.txt
( time timeout 1700 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r02h_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02h_pytest_gpu.log
tail -22 gpurun_out/r02h_pytest_gpu.log
