#!/usr/bin/env bash
# parity suite + kernel timings of every BASELINE shape, for this build and for tuning variants: bash tools/gpu_parity_timings.sh <tag> "<variants>"
set -u
TAG="${1:-r2a}"; VARS="${2:-}"
O=gpurun_out/$TAG; mkdir -p "$O"
timeout 1200 python -m pytest tests -m gpu -x -q > "$O/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$O/pytest_gpu.log"
tail -15 "$O/pytest_gpu.log"
run() { timeout 300 python tools/run_kernel.py "$@" --launches 5 2>&1 | tee -a "$O/run_kernel.log"; }
for cfg in "--config C2 --sites 1000000" "--config C3 --sites 100000" "--config C5 --sites 200000" "--config C5 --sites 200000 --abs-mode 1" "--config C4 --sites 9472"; do
  echo "default: $cfg" | tee -a "$O/run_kernel.log"
  run $cfg
  for v in $VARS; do
    echo "variant $v:" | tee -a "$O/run_kernel.log"
    BASEVAR_B200_LIB=$PWD/basevar_b200/variants/libbv_$v.so run $cfg
  done
done
