#!/usr/bin/env bash
# ncu full capture of the site kernel on one config: bash tools/gpu_prof.sh <tag> <config> <sites>
set -u
TAG="$1"; CFG="${2:-C2}"; SITES="${3:-1000000}"
O=gpurun_out/$TAG; mkdir -p "$O"
timeout 600 python -m pytest tests -m gpu -x -q > "$O/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$O/pytest_gpu.log"
tail -5 "$O/pytest_gpu.log"
timeout 300 python tools/run_kernel.py --config $CFG --sites $SITES --launches 5 2>&1 | tee "$O/run_kernel_$CFG.log"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bv_.*_kernel -s 5 -c 4 -f -o "$O/prof_$CFG" \
    python tools/run_kernel.py --config $CFG --sites $SITES --launches 3 > "$O/ncu_full_$CFG.log" 2>&1
tail -3 "$O/ncu_full_$CFG.log"
