#!/usr/bin/env bash
# round 2, call 5: full GPU suite (device-side slab counts, slab appends, device MapGridEmit query) + 1-GPU bench + racecheck of the list kernel
set -uo pipefail
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r02e_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02e_pytest_gpu.log
tail -22 gpurun_out/r02e_pytest_gpu.log
( time timeout 600 python bench.py --steps 50 --warmup 20 --no-extra-configs ) > gpurun_out/r02e_bench.log 2>&1
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02e_bench.log") if l.startswith("{")][-1]); print(round(d["ms_per_step"],4), {k: round(v,4) for k,v in d["roofline"]["phases_ms_per_step"].items()}, "e2e ms", round(d["e2e"]["ms_per_step"],3), "parity", d["parity"]["ok"])
except Exception as ex:
    print("bench FAILED", ex, open("gpurun_out/r02e_bench.log").read()[-1500:])
PY
# racecheck of the warp-synchronous shared-memory staging of the list build (small scene: the tool slows kernels ~100x)
( time timeout 900 compute-sanitizer --tool racecheck --kernel-regex kns=k_cell_lists python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02e_racecheck.log 2>&1
tail -5 gpurun_out/r02e_racecheck.log
