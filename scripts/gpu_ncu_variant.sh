#!/usr/bin/env bash
# `ncu --set full` of the sweeps of one variant library: gpu_ncu_variant.sh NAME [regex]
set -uo pipefail
mkdir -p gpurun_out
NAME=$1; PAT="${2:-k_force_np_predict|k_pressure}"
BBX_LIB=$PWD/bubbles_b200/lib/variants/libbbx_${NAME}.so timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$PAT" -s 36 -c 3 \
    -f -o gpurun_out/var_${NAME} python bench.py --steps 2 --warmup 12 --repeats 1 --no-parity --no-extra-configs --developed-substeps 0 --no-cpu-baseline --e2e-steps 1 > gpurun_out/var_${NAME}_ncu.log 2>&1
tail -c 300 gpurun_out/var_${NAME}_ncu.log
