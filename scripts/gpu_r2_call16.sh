#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_slabs.py tests/test_gpu_colliders_moving.py -m gpu -q -x --durations=5 ) > gpurun_out/r02r_pytest_slabs.log 2>&1
grep -v "^$" gpurun_out/r02r_pytest_slabs.log | tail -40 | cut -c1-250
