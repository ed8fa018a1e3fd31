#!/usr/bin/env bash
# round 2, call 2: GPU test suite on the v8 list kernel, A/B of list-kernel variants, ncu of one sub-step
set -uo pipefail
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 ) > gpurun_out/r02c_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02c_pytest_gpu.log
tail -25 gpurun_out/r02c_pytest_gpu.log
for so in v7 v8 v9 v9p896 v9l1alloc; do
  BBX_LIB=$PWD/bubbles_b200/lib/variants/libbbx_${so}.so timeout 300 python bench.py --steps 30 --warmup 20 --repeats 3 --no-cpu-baseline --no-extra-configs --developed-substeps 0 --e2e-steps 2 > gpurun_out/r02c_var_${so}.json 2> gpurun_out/r02c_var_${so}.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02c_var_${so}.json") if l.startswith("{")][-1]); print("${so}", round(d["ms_per_step"],4), {k: round(v,4) for k,v in d["roofline"]["phases_ms_per_step"].items()}, d["stats"]["exact_passes"], d["stats"].get("unstaged_tiles"), "parity", d["parity"]["ok"], d["parity"]["lists_bit_exact"])
except Exception as ex:
    print("${so} FAILED", ex, open("gpurun_out/r02c_var_${so}.err").read()[-800:])
PY
done
# one sub-step of the in-tree library under ncu --set full (after 30 warm-up sub-steps): list build + the three sweeps + grid kernels
timeout 900 ncu --set full --clock-control none --import-source on \
    --kernel-name "regex:k_fill_incremental|k_cell_lists_density|k_force_np_predict|k_pressure|k_hash_count|k_scan_cells" \
    --launch-skip 184 --launch-count 7 -f -o gpurun_out/r02c_full \
    python bench.py --steps 2 --warmup 30 --repeats 1 --no-cpu-baseline --no-extra-configs --no-parity --developed-substeps 0 --e2e-steps 1 > gpurun_out/r02c_ncu_full.log 2>&1
tail -2 gpurun_out/r02c_ncu_full.log
ls -la gpurun_out | tail -12
