#!/usr/bin/env bash
# build a kernel variant of libbbx.so for A/B runs on the GPU box: build_variant.sh NAME [-DFLAG ...]
set -euo pipefail
NAME=$1; shift
mkdir -p bubbles_b200/lib/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
    -o bubbles_b200/lib/variants/libbbx_${NAME}.so bubbles_b200/csrc/bbx_engine.cu -ldl -lpthread
