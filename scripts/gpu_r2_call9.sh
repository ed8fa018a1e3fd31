#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
( time timeout 800 compute-sanitizer --tool initcheck --print-limit 40 python -m pytest tests/test_gpu_colliders_moving.py -m gpu -q -x -k "slab_cut" ) > gpurun_out/r02k_initcheck.log 2>&1
grep -A12 "Uninitialized" gpurun_out/r02k_initcheck.log | grep -v "^=========         in \|cuda\|libcuda" | head -80 | cut -c1-200
tail -5 gpurun_out/r02k_initcheck.log
