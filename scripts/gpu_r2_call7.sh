#!/usr/bin/env bash
# round 2, call 7: thread-per-particle list build (v11): GPU suite + bench, A/B against the v7 list kernel
set -uo pipefail
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/r02i_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02i_pytest_gpu.log
tail -15 gpurun_out/r02i_pytest_gpu.log
( time timeout 600 python bench.py --steps 50 --warmup 20 --no-extra-configs ) > gpurun_out/r02i_bench.log 2>&1
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02i_bench.log") if l.startswith("{")][-1]); print(round(d["ms_per_step"],4), {k: round(v,4) for k,v in d["roofline"]["phases_ms_per_step"].items()}, "e2e ms", round(d["e2e"]["ms_per_step"],3), "parity", d["parity"]["ok"], d["stats"])
except Exception as ex:
    print("bench FAILED", ex, open("gpurun_out/r02i_bench.log").read()[-1500:])
PY
./scripts/gpu_variants.sh v7
