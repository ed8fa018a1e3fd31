#!/usr/bin/env bash
# compute-sanitizer passes over the small parity scene (3 k particles, 3 sub-steps = __graft_entry__.smoke()):
#   memcheck  -- out-of-bounds / misaligned accesses in every kernel of the sub-step
#   racecheck -- shared-memory hazards of the warp-synchronous code (list build: per-lane list buffers, run buffers filled by
#                cp.async.bulk behind an mbarrier; k_pressure: staged tile)
#   synccheck -- barrier / mbarrier misuse
# Run on a GPU box: gpurun --timeout 900 -- 'bash scripts/gpu_sanitize.sh'; logs land in gpurun_out/sanitize_*.log
set -uo pipefail
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 280 compute-sanitizer --tool $tool --error-exitcode 9 python __graft_entry__.py --smoke > gpurun_out/sanitize_${tool}.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/sanitize_${tool}.log
done
