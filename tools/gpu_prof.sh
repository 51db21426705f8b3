#!/usr/bin/env bash
# ncu full capture of chosen kernels: bash tools/gpu_prof.sh <tag> <regex> "<cfg sites [abs]>;..."
set -u
TAG="$1"; RX="$2"; CFGS="${3:-C5 200000 0}"
O=gpurun_out/$TAG; mkdir -p "$O"
IFS=';' read -ra CL <<< "$CFGS"
for cfg in "${CL[@]}"; do
  set -- $cfg
  AB="${3:-0}"
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$RX" -s 2 -c 2 -f -o "$O/prof_$1_abs$AB" \
      python tools/run_kernel.py --config $1 --sites $2 --abs-mode $AB --launches 2 > "$O/ncu_$1_abs$AB.log" 2>&1
  tail -2 "$O/ncu_$1_abs$AB.log"
done
ls -la $O
