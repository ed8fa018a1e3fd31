#!/usr/bin/env bash
# A/B on the GPU box: one short bench per variant library in bubbles_b200/lib/variants (phase times + the parity gate)
set -uo pipefail
mkdir -p gpurun_out
for so in "$@"; do
  BBX_LIB=$PWD/bubbles_b200/lib/variants/libbbx_${so}.so timeout 300 python bench.py --steps 50 --warmup 20 --repeats 3 --no-cpu-baseline --no-extra-configs --e2e-steps 2 2>&1 | tail -1 > gpurun_out/var_${so}.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/var_${so}.json")); print("${so}", round(d["ms_per_step"],4), {k: round(v,4) for k,v in d["roofline"]["phases_ms_per_step"].items()}, "parity", d["parity"]["ok"], "developed", round(d["developed"]["ms_per_step"],4), d["developed"]["parity"]["ok"], d["stats"]["exact_passes"], d["stats"].get("unstaged_tiles"))
except Exception as ex:
    print("${so} FAILED", ex, open("gpurun_out/var_${so}.json").read()[-500:])
PY
done
