#!/usr/bin/env bash
# round 2 evidence call: whole GPU suite, the driver's default bench line, the ncu launch list of the same command,
# one `ncu --set full` capture of every kernel of one PCISPH sub-step and of one SPH sub-step
set -uo pipefail
T=r02x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/${T}_gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -q -x --durations=8 ) > gpurun_out/${T}_pytest_gpu.log 2>&1
grep -v "^$" gpurun_out/${T}_pytest_gpu.log | tail -14 | cut -c1-200
( time timeout 900 python bench.py ) > gpurun_out/${T}_bench_1gpu.log 2>&1
tail -c 1500 gpurun_out/${T}_bench_1gpu.log
SHORT="--steps 2 --warmup 12 --repeats 1 --no-parity --no-extra-configs --developed-substeps 0 --no-cpu-baseline --e2e-steps 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py $SHORT > gpurun_out/${T}_bench_ncu.log 2>&1
tail -c 300 gpurun_out/${T}_bench_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 150 --launch-count 16 -f -o gpurun_out/${T}_full \
    python bench.py $SHORT > gpurun_out/${T}_ncu_full.log 2>&1
tail -c 300 gpurun_out/${T}_ncu_full.log
timeout 600 ncu --set full --clock-control none --launch-skip 100 --launch-count 10 -f -o gpurun_out/${T}_full_sph \
    python bench.py --solver sph $SHORT > gpurun_out/${T}_ncu_full_sph.log 2>&1
tail -c 300 gpurun_out/${T}_ncu_full_sph.log
ls -la gpurun_out
