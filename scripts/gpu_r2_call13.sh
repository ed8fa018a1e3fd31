#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 ) > gpurun_out/r02n_pytest_gpu.log 2>&1
grep -v "^$" gpurun_out/r02n_pytest_gpu.log | tail -12 | cut -c1-220
