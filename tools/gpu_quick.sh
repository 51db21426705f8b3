#!/usr/bin/env bash
# Quick GPU session: parity tests + kernel timings of tuning variants.  bash tools/gpu_quick.sh <tag> "<variants>"
set -u
TAG="${1:-q}"; VARS="${2:-}"
O=gpurun_out/$TAG; mkdir -p "$O"
timeout 900 python -m pytest tests -m gpu -x -q > "$O/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$O/pytest_gpu.log"
tail -15 "$O/pytest_gpu.log"
for cfg in "C2 1000000" "C3 100000" "C5 200000"; do
  set -- $cfg
  timeout 300 python tools/run_kernel.py --config $1 --sites $2 --launches 5 2>&1 | tee -a "$O/run_kernel.log"
  for v in $VARS; do
    echo "variant $v:" | tee -a "$O/run_kernel.log"
    BASEVAR_B200_LIB=$PWD/basevar_b200/variants/libbv_$v.so timeout 300 python tools/run_kernel.py --config $1 --sites $2 --launches 5 2>&1 | tee -a "$O/run_kernel.log"
  done
done
