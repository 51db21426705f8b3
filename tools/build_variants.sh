#!/usr/bin/env bash
# Tuning builds of the CUDA library.  Each argument word is name:DEF[,DEF...]:
#   tools/build_variants.sh "c24:BV_COUNT_WARPS=24 c32s2:BV_COUNT_WARPS=32,BV_COUNT_STAGES=2"
# -> basevar_b200/variants/libbv_<name>.so
set -e
cd "$(dirname "$0")/../basevar_b200"
mkdir -p variants
for v in $1; do
  name=${v%%:*}; defs=$(echo "${v#*:}" | sed 's/,/ -D/g; s/^/-D/')
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -shared \
       $defs -o "variants/libbv_$name.so" csrc/bv_api.cu csrc/bv_encode16.cpp
done
ls variants
