#!/usr/bin/env bash
# The product on the GPU box: CLI parity tests, C1 and a 1,000-BAM cohort against the reference command: bash tools/gpu_cli.sh <tag>
set -u
TAG="${1:-cli2}"
O=gpurun_out/$TAG; mkdir -p "$O"
timeout 600 python -m pytest tests/test_pileup.py tests/test_host_cpp.py -m gpu -x -q > "$O/pytest_cli.log" 2>&1; tail -3 "$O/pytest_cli.log"
timeout 600 python tools/cli_bench.py --c1 bench_data/c1 --threads 4 --work /tmp/bv_cli_c1 > "$O/cli_c1.json" 2> "$O/cli_c1.err"; echo "c1 rc=$?"; cat "$O/cli_c1.json"
# where the start-up goes: the same command on an empty region of the small fixture
( time basevar_b200/bin/basevar basetype -R tests/golden/range/ce.fa.gz -I tests/golden/range/range.bam -r CHROMOSOME_I:1-10 --output-vcf /tmp/e.vcf --output-cvg /tmp/e.cvg ) > "$O/startup.txt" 2>&1; tail -4 "$O/startup.txt"
timeout 1500 python tools/cli_bench.py --bams 1000 --mb 10 --no-reference --threads $(nproc) --work /tmp/bv_cli_big > "$O/cli_cohort.json" 2> "$O/cli_cohort.err"; echo "cohort rc=$?"; cat "$O/cli_cohort.json"; tail -5 "$O/cli_cohort.err"
