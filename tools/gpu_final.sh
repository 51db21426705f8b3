#!/usr/bin/env bash
# The round's last run on one B200, on the final binary: smoke, both bench arms, per-shape kernel timings (the parity suite, ncu captures and
# compute-sanitizer of this code are those of tools/gpu_r02z.sh; what changed since is the start state of the right Fisher walk).
set -u
O=gpurun_out/${1:-final}; mkdir -p "$O"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$O/smoke.log" 2>&1; tail -1 "$O/smoke.log"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > "$O/bench_reference.json" 2> "$O/bench_reference.err"
timeout 900 python bench.py > "$O/bench.json" 2> "$O/bench.err"; tail -2 "$O/bench.err"
python tools/bench_show.py "$O/bench.json"
for cfg in "C2 1000000 0" "C3 100000 0" "C5 200000 0" "C5 1000000 0" "C5 200000 1" "C4 9472 0"; do
  set -- $cfg
  timeout 300 python tools/run_kernel.py --config $1 --sites $2 --abs-mode $3 --launches 5 2>&1 | tee -a "$O/run_kernel.log"
done
