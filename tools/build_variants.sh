#!/usr/bin/env bash
# Tuning builds of the CUDA library: tools/build_variants.sh "28 24 20" -> basevar_b200/variants/libbv_w<N>.so
set -e
cd "$(dirname "$0")/../basevar_b200"
mkdir -p variants
for w in $1; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -shared \
       -DBV_WARPS=$w ${BV_EXTRA_FLAGS:-} -o variants/libbv_w$w.so csrc/bv_api.cu
done
ls -la variants
