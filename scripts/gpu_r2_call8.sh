#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 300 python scripts/debug/lists_debug.py > gpurun_out/r02j_lists_debug.log 2>&1
cut -c1-300 gpurun_out/r02j_lists_debug.log | head -30
