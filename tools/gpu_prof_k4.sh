#!/usr/bin/env bash
# ncu full capture of the K4 kernels (histogram, EM tasks, decision) on several shapes: bash tools/gpu_prof_k4.sh <tag> [regex]
set -u
TAG="$1"; RX="${2:-bv_(hist|em_task|decide)_kernel}"
O=gpurun_out/$TAG; mkdir -p "$O"
for cfg in "C5 200000" "C2 1000000" "C4 9472"; do
  set -- $cfg
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$RX" -s 3 -c 3 -f -o "$O/prof_$1" \
      python tools/run_kernel.py --config $1 --sites $2 --launches 2 > "$O/ncu_$1.log" 2>&1
  tail -2 "$O/ncu_$1.log"
done
ls -la $O
