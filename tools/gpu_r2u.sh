#!/usr/bin/env bash
# parity suite, smoke, timings (+ variants), both bench arms
set -u
TAG="${1:-r2u}"; VARS="${2:-}"
O=gpurun_out/$TAG; mkdir -p "$O"
nvidia-smi > "$O/nvidia-smi.txt" 2>&1
bash tools/gpu_r2a.sh "$TAG" "$VARS" > "$O/r2a.log" 2>&1; tail -4 "$O/pytest_gpu.log"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$O/smoke.log" 2>&1; tail -2 "$O/smoke.log"
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > "$O/bench_reference.json" 2> "$O/bench_reference.err" ) 2>&1 | grep real
( time timeout 900 python bench.py > "$O/bench.json" 2> "$O/bench.err" ) 2>&1 | grep real
tail -3 "$O/bench.err"
python tools/bench_show.py "$O/bench.json"
python tools/variants_table.py "$O/run_kernel.log"
