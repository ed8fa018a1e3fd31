#!/usr/bin/env bash
# Run on an N-GPU box (gpurun --gpus N): NCCL slab check + bench with the halo push (default), and the bench
# again with BBX_P2P=0 (send / recv per phase).  usage: gpu_multi.sh N TAG [extra bench args...]
set -uo pipefail
N=${1:-2}; TAG=${2:-multi}; shift 2 || true
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "${SKIP_CHECK:-0}" != "1" ]; then
  timeout 240 $TR scripts/slab_nccl_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -1 | tee gpurun_out/${TAG}_check_${N}gpu.json
fi
for P in ${MODES:-1 0}; do
  BBX_P2P=$P timeout 600 $TR bench.py --gpus $N --steps ${STEPS:-50} --warmup ${WARMUP:-10} "$@" 2>&1 | grep "^{" | tail -1 > gpurun_out/${TAG}_bench_${N}gpu_p2p$P.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_${N}gpu_p2p$P.json")); print("p2p=$P", "%.3e"%d["value"], round(d["ms_per_step"],4), {k: round(v,4) for k,v in d["roofline"]["phases_ms_per_step"].items()}, "e2e %.3e"%d["e2e"]["value"], d["roofline"]["whole_step"]["frac"])
except Exception as ex:
    print("bench p2p=$P FAILED", ex)
PY
done
