#!/usr/bin/env bash
# Run on the B200 box (under gpurun): GPU parity tests, one bench line, the ncu launch list of the same bench
# command, and one `ncu --set full` capture of the hot kernels of one sub-step.  TAG names the outputs.
set -uo pipefail
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
fi
( time timeout 900 python bench.py --steps 50 --warmup 10 ${BENCH_ARGS:-} ) > gpurun_out/${TAG}_bench.log 2>&1
tail -3 gpurun_out/${TAG}_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_bench_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_bench_ncu.log
# hot kernels of the 12th sub-step (after the warm-up): skip the first 11 launches of each
timeout 900 ncu --set full --clock-control none --import-source on \
    --kernel-name "regex:k_fill_incremental|k_cell_lists_density|k_force_np_predict|k_pressure|k_hash_count|k_scan_cells" \
    --launch-skip 65 --launch-count 7 -f -o gpurun_out/${TAG}_full \
    python bench.py --steps 2 --warmup 12 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out
