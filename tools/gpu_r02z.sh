#!/usr/bin/env bash
# Round-2 evidence run on one B200: parity suite (flip lists written), smoke, both bench arms, per-shape kernel timings, the ncu
# launch list of the bench command, full ncu captures of one whole step on C2 / C5 / the C4 tile and of K0, compute-sanitizer.
#   bash tools/gpu_r02z.sh <tag>
set -u
TAG="${1:-r02z}"
O=gpurun_out/$TAG; mkdir -p "$O"
nvidia-smi > "$O/nvidia-smi.txt" 2>&1
BV_WRITE_FLIPS=1 timeout 1500 python -m pytest tests -m gpu -q > "$O/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$O/pytest_gpu.log"
tail -4 "$O/pytest_gpu.log"; cp gpurun_out/flips_observed.json "$O/" 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$O/smoke.log" 2>&1; tail -1 "$O/smoke.log"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > "$O/bench_reference.json" 2> "$O/bench_reference.err"
timeout 900 python bench.py > "$O/bench.json" 2> "$O/bench.err"; tail -2 "$O/bench.err"
python tools/bench_show.py "$O/bench.json"
for cfg in "C2 1000000 0" "C3 100000 0" "C5 200000 0" "C5 1000000 0" "C5 200000 1" "C4 9472 0"; do
  set -- $cfg
  timeout 300 python tools/run_kernel.py --config $1 --sites $2 --abs-mode $3 --launches 5 2>&1 | tee -a "$O/run_kernel.log"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file "$O/launches.csv" \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --e2e-steps 1 > "$O/bench_under_ncu.log" 2>&1
python tools/ncu_launches.py "$O/launches.csv" > "$O/launches_by_kernel.txt" 2>&1; head -20 "$O/launches_by_kernel.txt"
for cfg in "C2 1000000" "C5 200000" "C4 9472"; do
  set -- $cfg
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:bv_(count|scalar|bound|hist|em_task|fisher)_kernel" -s 7 -c 7 -f \
      -o "$O/prof_$1_step" python tools/run_kernel.py --config $1 --sites $2 --launches 2 > "$O/ncu_$1.log" 2>&1
  tail -1 "$O/ncu_$1.log"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bv_expand_kernel -s 3 -c 1 -f -o "$O/prof_K0_C4" \
    python tools/e2e_sweep.py --config C4 --sites 18944 --u16 --tiles 9472 --slots 4 --reps 1 > "$O/ncu_K0_C4.log" 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bv_expand_kernel -s 9 -c 1 -f -o "$O/prof_K0_C2" \
    python tools/e2e_sweep.py --config C2 --sites 1000000 --u16 --tiles 131072 --slots 3 --reps 1 > "$O/ncu_K0_C2.log" 2>&1
{
  echo "== memcheck: sparse tiles (edge cases, malformed input, 100,000-sample rows, chunk edges), parity (golden sites, fuzz, ties)"
  timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_parity.py -m gpu -q -x \
      -k "edge or malformed or tiny or chunk or 100000 or golden_sites or ties or C1-like" 2>&1 | tail -6
  echo "== racecheck: the same selection without the fuzz"
  timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_parity.py tests/test_gpu_calls.py -m gpu -q -x \
      -k "edge or tiny or chunk-edge or 100000 or golden_sites or ties or C1-like or fixture" 2>&1 | tail -6
} > "$O/compute_sanitizer.txt" 2>&1
cat "$O/compute_sanitizer.txt"
ls -la "$O"
