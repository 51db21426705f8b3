#!/usr/bin/env bash
# Round-end GPU session: parity tests, smoke, bench (both arms), per-shape kernel timings, e2e sweep, the ncu launch list of
# the bench command and full captures (C2 step kernels, C5 EM kernel, K0 expand kernel).  bash tools/gpu_final.sh <tag>
set -u
TAG="${1:-final}"
O=gpurun_out/$TAG; mkdir -p "$O"
nvidia-smi > "$O/nvidia-smi.txt" 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > "$O/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$O/pytest_gpu.log"
tail -4 "$O/pytest_gpu.log"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$O/smoke.log" 2>&1; tail -2 "$O/smoke.log"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > "$O/bench_reference.json" 2> "$O/bench_reference.err"
timeout 900 python bench.py > "$O/bench.json" 2> "$O/bench.err"; cat "$O/bench.json"; tail -3 "$O/bench.err"
for cfg in "C2 1000000" "C3 100000" "C5 200000"; do
  set -- $cfg
  timeout 300 python tools/run_kernel.py --config $1 --sites $2 --launches 5 2>&1 | tee -a "$O/run_kernel.log"
done
timeout 300 python tools/run_kernel.py --config C5 --sites 200000 --launches 5 --abs-mode 1 2>&1 | tee -a "$O/run_kernel.log"
timeout 300 python tools/e2e_sweep.py --config C2 --sites 1000000 --tiles 32768,131072 --slots 3 2>&1 | tee "$O/e2e_sweep.log"
timeout 300 python tools/e2e_sweep.py --config C2 --sites 1000000 --u16 --tiles 32768,131072 --slots 3 2>&1 | tee -a "$O/e2e_sweep.log"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$O/launches.csv" \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > "$O/bench_under_ncu.log" 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bv_.*_kernel -s 5 -c 4 -f -o "$O/prof_C2" \
    python tools/run_kernel.py --config C2 --sites 1000000 --launches 3 > "$O/ncu_full_C2.log" 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bv_em_kernel -s 2 -c 1 -f -o "$O/prof_C5_em" \
    python tools/run_kernel.py --config C5 --sites 200000 --launches 2 > "$O/ncu_full_C5.log" 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bv_expand_kernel -s 9 -c 1 -f -o "$O/prof_K0" \
    python tools/e2e_sweep.py --config C2 --sites 1000000 --u16 --tiles 131072 --slots 3 --reps 1 > "$O/ncu_full_K0.log" 2>&1
ls -la "$O"
