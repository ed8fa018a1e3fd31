#!/usr/bin/env bash
# round 2, multi-GPU call: N = $1 ranks -- NCCL + IPC slab path bit-identical to the single domain, then the bench line
set -uo pipefail
N=${1:-2}
TAG=${2:-r02f}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/slab_nccl_check.py ) > gpurun_out/${TAG}_nccl_check_${N}gpu.log 2>&1
grep '^{' gpurun_out/${TAG}_nccl_check_${N}gpu.log | tail -1 | cut -c1-600
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 50 --warmup 20 ) > gpurun_out/${TAG}_bench_${N}gpu.log 2>&1
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${TAG}_bench_${N}gpu.log") if l.startswith("{")][-1])
    print("N=$N", round(d["ms_per_step"],4), "value %.3e" % d["value"], {k: round(v,4) for k,v in d["roofline"]["phases_ms_per_step"].items()}, "e2e %.3e" % d["e2e"]["value"], "parity", d["parity"] and d["parity"]["ok"], "developed", d["developed"] and round(d["developed"]["ms_per_step"],4))
    print("configs", {k: (v.get("ms_per_step"), v.get("value"), v.get("whole_step_frac"), v.get("failed")) for k,v in d["configs"].items()})
except Exception as ex:
    print("bench FAILED", ex, open("gpurun_out/${TAG}_bench_${N}gpu.log").read()[-2500:])
PY
if [ "$N" = "2" ]; then
( time timeout 600 python bench.py --steps 50 --warmup 20 --no-extra-configs --no-cpu-baseline ) > gpurun_out/${TAG}_bench_1gpu.log 2>&1
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${TAG}_bench_1gpu.log") if l.startswith("{")][-1]); print("N=1", round(d["ms_per_step"],4), {k: round(v,4) for k,v in d["roofline"]["phases_ms_per_step"].items()}, "parity", d["parity"]["ok"])
except Exception as ex:
    print("bench1 FAILED", ex)
PY
fi
