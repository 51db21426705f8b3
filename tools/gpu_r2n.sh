#!/usr/bin/env bash
set -u
TAG="${1:-r2n}"
O=gpurun_out/$TAG; mkdir -p "$O"
timeout 1200 python -m pytest tests -m gpu -x -q > "$O/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$O/pytest_gpu.log"; tail -4 "$O/pytest_gpu.log"
timeout 600 python tools/cli_bench.py --c1 bench_data/c1 --threads 4 --work /tmp/bv_cli_c1 > "$O/cli_c1.json" 2> "$O/cli_c1.err"; echo "c1 rc=$?"; cat "$O/cli_c1.json"
timeout 1500 python tools/cli_bench.py --bams 1000 --mb 10 --no-reference --threads $(nproc) --work /tmp/bv_cli_big > "$O/cli_cohort.json" 2> "$O/cli_cohort.err"; echo "cohort rc=$?"; cat "$O/cli_cohort.json"; tail -5 "$O/cli_cohort.err"
