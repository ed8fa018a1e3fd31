#!/usr/bin/env bash
# round 2: whole GPU suite + the default bench line (what the driver runs at round end)
set -uo pipefail
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x --durations=8 ) > gpurun_out/r02w_pytest_gpu.log 2>&1
grep -v "^$" gpurun_out/r02w_pytest_gpu.log | tail -25 | cut -c1-250
( time timeout 900 python bench.py ) > gpurun_out/r02w_bench_1gpu.log 2>&1
tail -c 3000 gpurun_out/r02w_bench_1gpu.log
