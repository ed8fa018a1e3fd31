#!/usr/bin/env bash
# ncu --set full of the list build v12 (one launch after warm-up)
set -uo pipefail
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_lists_density_tp" -s 25 -c 1 \
    -f -o gpurun_out/r02o_tp12 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extra-configs > gpurun_out/r02o_ncu.log 2>&1
tail -2 gpurun_out/r02o_ncu.log | cut -c1-200
