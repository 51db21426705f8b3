#!/usr/bin/env bash
# Multi-GPU checks (under gpurun --gpus N): the bench at N ranks (C4 shards), the reference arm, the sharded CLI test.
set -u
TAG="${1:-multi}"; N="${2:-2}"
O=gpurun_out/$TAG; mkdir -p "$O"
nvidia-smi -L > "$O/gpus.txt"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > "$O/bench_n$N.json" 2> "$O/bench_n$N.err"; echo "bench N=$N rc=$?"; tail -c 800 "$O/bench_n$N.err"
python tools/bench_show.py "$O/bench_n$N.json"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > "$O/bench_ref_n$N.json" 2> "$O/bench_ref_n$N.err"; echo "reference N=$N rc=$?"; cut -c1-700 "$O/bench_ref_n$N.json"
timeout 600 python -m pytest tests/test_pileup.py -m gpu -q -k "several_gpus" > "$O/pytest_multi.log" 2>&1; tail -3 "$O/pytest_multi.log"
