#!/usr/bin/env bash
set -u
TAG="${1:-r2g}"; VARS="${2:-}"
O=gpurun_out/$TAG; mkdir -p "$O"
bash tools/gpu_r2a.sh "$TAG" "$VARS"
timeout 900 python bench.py > "$O/bench.json" 2> "$O/bench.err"; echo "bench rc=$?"; tail -c 1500 "$O/bench.err"
for sl in 3 6; do
  timeout 600 python bench.py --slots $sl --no-configs --no-cpu-baseline --no-e2e-extra > "$O/bench_slots$sl.json" 2> "$O/bench_slots$sl.err"; echo "bench slots $sl rc=$?"
done
python tools/bench_show.py "$O/bench.json" "$O/bench_slots3.json" "$O/bench_slots6.json"
