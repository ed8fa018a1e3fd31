#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
./scripts/gpu_variants.sh u8 r1 u8r1 r3
