#!/usr/bin/env bash
# One `ncu --set full` capture of each hot kernel (one launch each, after warm-up) on a short bench run.
set -uo pipefail
mkdir -p gpurun_out
PAT="${1:-k_cell_lists_density|k_cell_force_np_predict|k_cell_pressure|k_cell_pressure_force}"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$PAT" -s ${3:-25} -c ${2:-6} \
    -f -o gpurun_out/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
