#!/usr/bin/env bash
# Session: sparse-tile tests first, full GPU suite, bench, C5 ncu capture of the finish kernels.
set -u
TAG="${1:-s1}"
O=gpurun_out/$TAG; mkdir -p "$O"
timeout 600 python -m pytest tests/test_gpu_sparse.py -x -q > "$O/pytest_sparse.log" 2>&1; echo "rc=$?" >> "$O/pytest_sparse.log"
tail -25 "$O/pytest_sparse.log"
timeout 900 python -m pytest tests -m gpu -x -q > "$O/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$O/pytest_gpu.log"
tail -5 "$O/pytest_gpu.log"
timeout 600 python bench.py > "$O/bench.json" 2> "$O/bench.err"; cat "$O/bench.json"; tail -5 "$O/bench.err"
timeout 300 python tools/run_kernel.py --config C5 --sites 200000 --launches 5 2>&1 | tee "$O/run_kernel_C5.log"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bv_.*_kernel -s 4 -c 4 -f -o "$O/prof_C5" \
    python tools/run_kernel.py --config C5 --sites 200000 --launches 2 > "$O/ncu_full_C5.log" 2>&1
tail -3 "$O/ncu_full_C5.log"
ls -la "$O"
