#!/usr/bin/env bash
# round 2: parity + timings of every shape, then full ncu captures of one whole step (all five kernels) on C2 and C5
set -u
TAG="${1:-r2r}"; VARS="${2:-}"
O=gpurun_out/$TAG; mkdir -p "$O"
bash tools/gpu_r2a.sh "$TAG" "$VARS"
for cfg in "C2 1000000 0" "C5 200000 0"; do
  set -- $cfg
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:bv_(count|scalar|bound|hist|em_task)_kernel" -s 5 -c 5 -f \
      -o "$O/prof_$1_step" python tools/run_kernel.py --config $1 --sites $2 --abs-mode $3 --launches 2 > "$O/ncu_$1.log" 2>&1
  tail -1 "$O/ncu_$1.log"
done
ls -la "$O"
