#!/usr/bin/env bash
set -u
TAG="${1:-r2i}"
O=gpurun_out/$TAG; mkdir -p "$O"
bash tools/gpu_r2a.sh "$TAG" ""
timeout 900 python bench.py --no-configs > "$O/bench.json" 2> "$O/bench.err"; echo "bench rc=$?"; tail -c 1500 "$O/bench.err"
python tools/bench_show.py "$O/bench.json"
