#!/usr/bin/env bash
# The product against the reference command on the GPU box: bash tools/gpu_cli.sh <tag>
set -u
TAG="${1:-cli}"
O=gpurun_out/$TAG; mkdir -p "$O"
nproc > "$O/nproc.txt"
timeout 600 python tools/cli_bench.py --c1 bench_data/c1 --threads 4 --work /tmp/bv_cli_c1 > "$O/cli_c1.json" 2> "$O/cli_c1.err"; echo "c1 rc=$?"; cat "$O/cli_c1.json"
timeout 1500 python tools/cli_bench.py --bams 1000 --mb 10 --ref-mb 0.4 --threads $(nproc) --work /tmp/bv_cli_big > "$O/cli_cohort.json" 2> "$O/cli_cohort.err"; echo "cohort rc=$?"; cat "$O/cli_cohort.json"; tail -5 "$O/cli_cohort.err"
