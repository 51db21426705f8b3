#!/usr/bin/env bash
# parity suite, smoke and both bench arms on one GPU: bash tools/gpu_bench_only.sh <tag>
set -u
TAG="${1:-bench}"
O=gpurun_out/$TAG; mkdir -p "$O"
timeout 1200 python -m pytest tests -m gpu -q > "$O/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$O/pytest_gpu.log"; tail -3 "$O/pytest_gpu.log"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$O/smoke.log" 2>&1; tail -1 "$O/smoke.log"
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > "$O/bench_reference.json" 2> "$O/bench_reference.err" ) 2>&1 | grep real
( time timeout 900 python bench.py > "$O/bench.json" 2> "$O/bench.err" ) 2>&1 | grep real
tail -2 "$O/bench.err"; python tools/bench_show.py "$O/bench.json"
