#!/usr/bin/env bash
# round 2, call 4: GPU tests on the device-side slab counts + list kernel v10 (warp-private bulk copies), A/B v7 / v10
set -uo pipefail
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r02d_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02d_pytest_gpu.log
tail -22 gpurun_out/r02d_pytest_gpu.log
for so in v7 v10; do
  BBX_LIB=$PWD/bubbles_b200/lib/variants/libbbx_${so}.so timeout 300 python bench.py --steps 30 --warmup 20 --repeats 3 --no-cpu-baseline --no-extra-configs --developed-substeps 0 --e2e-steps 2 > gpurun_out/r02d_var_${so}.json 2> gpurun_out/r02d_var_${so}.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02d_var_${so}.json") if l.startswith("{")][-1]); print("${so}", round(d["ms_per_step"],4), {k: round(v,4) for k,v in d["roofline"]["phases_ms_per_step"].items()}, d["stats"]["exact_passes"], d["stats"].get("unstaged_tiles"), "parity", d["parity"]["ok"], d["parity"]["lists_bit_exact"])
except Exception as ex:
    print("${so} FAILED", ex, open("gpurun_out/r02d_var_${so}.err").read()[-800:])
PY
done
timeout 900 ncu --set full --clock-control none --import-source on \
    --kernel-name "regex:k_cell_lists_density" \
    --launch-skip 27 --launch-count 1 -f -o gpurun_out/r02d_lists \
    python bench.py --steps 2 --warmup 30 --repeats 1 --no-cpu-baseline --no-extra-configs --no-parity --developed-substeps 0 --e2e-steps 1 > gpurun_out/r02d_ncu.log 2>&1
tail -2 gpurun_out/r02d_ncu.log
