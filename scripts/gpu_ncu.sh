#!/usr/bin/env bash
# Run on the B200 box (under gpurun): bench line + `ncu --set full` of the kernels matching REGEX in one
# sub-step after the warm-up.  usage: gpu_ncu.sh TAG REGEX PER_STEP [SETUP_LAUNCHES] [BENCH_ARGS...]
set -uo pipefail
TAG=$1; REGEX=$2; PER=$3; SETUP=${4:-0}; shift 4 || shift $#
mkdir -p gpurun_out
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline "$@" 2>&1 | tail -1 > gpurun_out/${TAG}_bench.json
python - <<PY
import json; d=json.load(open("gpurun_out/${TAG}_bench.json")); print(d["value"], d["ms_per_step"], d["roofline"]["phases_ms_per_step"], d["stats"], d["e2e"]["value"])
PY
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name "regex:${REGEX}" \
    --launch-skip $((SETUP + PER * 10)) --launch-count ${PER} -f -o gpurun_out/${TAG}_full \
    python bench.py --steps 2 --warmup 12 --no-cpu-baseline --e2e-steps 1 "$@" > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
