#!/usr/bin/env bash
# Run on the B200 box (under gpurun): GPU parity tests, one bench line, ncu launch list of the same bench command.
set -uo pipefail
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py --steps 50 --warmup 10 ) > gpurun_out/bench.log 2>&1
tail -3 gpurun_out/bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_ncu.log 2>&1
tail -2 gpurun_out/bench_ncu.log
