#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 300 python scripts/debug/lists_debug.py > gpurun_out/r02p_lists_debug.log 2>&1
cut -c1-200 gpurun_out/r02p_lists_debug.log | head -8
( time timeout 600 python bench.py --steps 50 --warmup 20 --no-extra-configs --no-cpu-baseline ) > gpurun_out/r02p_bench.log 2>&1
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02p_bench.log") if l.startswith("{")][-1]); print(round(d["ms_per_step"],4), {k: round(v,4) for k,v in d["roofline"]["phases_ms_per_step"].items()}, "e2e ms", round(d["e2e"]["ms_per_step"],3), "parity", d["parity"]["ok"], d["developed"]["parity"]["ok"], d["developed"]["ms_per_step"])
except Exception as ex:
    print("bench FAILED", ex, open("gpurun_out/r02p_bench.log").read()[-1500:])
PY
