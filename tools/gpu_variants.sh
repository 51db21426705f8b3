#!/usr/bin/env bash
# Kernel timings of tuning variants: bash tools/gpu_variants.sh <tag> "<variants>" "<cfg sites [samples]>;..."
set -u
TAG="$1"; VARS="$2"; CFGS="${3:-C2 1000000;C3 100000}"
O=gpurun_out/$TAG; mkdir -p "$O"
IFS=';' read -ra CL <<< "$CFGS"
for cfg in "${CL[@]}"; do
  set -- $cfg
  EXTRA=""; [ $# -ge 3 ] && EXTRA="--samples $3"
  echo "default:" | tee -a "$O/variants.log"
  timeout 300 python tools/run_kernel.py --config $1 --sites $2 $EXTRA --launches 5 2>&1 | tee -a "$O/variants.log"
  for v in $VARS; do
    echo "variant $v:" | tee -a "$O/variants.log"
    BASEVAR_B200_LIB=$PWD/basevar_b200/variants/libbv_$v.so timeout 300 python tools/run_kernel.py --config $1 --sites $2 $EXTRA --launches 5 2>&1 | tee -a "$O/variants.log"
  done
done
