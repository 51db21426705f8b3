#!/usr/bin/env bash
# One GPU-box session: parity tests, bench (both arms), ncu launch list and one full capture of the site kernel.
# Usage (under gpurun): bash tools/gpu_round.sh <tag>
set -u
TAG="${1:-run}"
O=gpurun_out/$TAG
mkdir -p "$O"
nvidia-smi > "$O/nvidia-smi.txt" 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > "$O/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$O/pytest_gpu.log"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > "$O/bench_reference.json" 2> "$O/bench_reference.err"
timeout 600 python bench.py > "$O/bench.json" 2> "$O/bench.err"
timeout 300 python tools/run_kernel.py --config C2 --sites 1000000 --launches 5 > "$O/run_kernel_C2.log" 2>&1
timeout 300 python tools/run_kernel.py --config C3 --sites 100000 --launches 5 > "$O/run_kernel_C3.log" 2>&1
timeout 300 python tools/run_kernel.py --config C5 --sites 200000 --launches 5 > "$O/run_kernel_C5.log" 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$O/launches.csv" \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > "$O/bench_under_ncu.log" 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bv_.*_kernel -s 5 -c 4 -f -o "$O/prof_C2" \
    python tools/run_kernel.py --config C2 --sites 1000000 --launches 3 > "$O/ncu_full.log" 2>&1
tail -3 "$O/pytest_gpu.log"; cat "$O/bench.json"; cat "$O/run_kernel_"*.log
