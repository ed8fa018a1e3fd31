#!/usr/bin/env bash
# round 2, call 1: the whole GPU test suite (incl. the new config-scale parity gate) + one bench line
set -uo pipefail
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r02a_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest_gpu.log
tail -30 gpurun_out/r02a_pytest_gpu.log
( time timeout 900 python bench.py --steps 50 --warmup 20 ) > gpurun_out/r02a_bench.log 2>&1
tail -c 6000 gpurun_out/r02a_bench.log
( time timeout 600 python bench.py --impl reference --steps 10 --warmup 2 ) > gpurun_out/r02a_bench_ref.log 2>&1
tail -c 1500 gpurun_out/r02a_bench_ref.log
