#!/usr/bin/env bash
# End-to-end legs on one GPU after a transport change: the tile timeline (BASEVAR_B200_TRACE=1), bench.py --config C4 over slots /
# tiles per step, the default bench line, the parity suite, the product timings.  bash tools/gpu_c4_e2e.sh <tag>
set -u
TAG="${1:-c4e2e}"
O=gpurun_out/$TAG; mkdir -p "$O"
BASEVAR_B200_TRACE=1 timeout 300 python tools/e2e_sweep.py --config C4 --sites 75776 --u16 --tiles 9472 --slots 4 --reps 2 > "$O/trace_c4.log" 2>&1; tail -14 "$O/trace_c4.log" | cut -c1-250
run() { n="$1"; shift; timeout 600 python bench.py --no-configs --no-cpu-baseline "$@" > "$O/bench_$n.json" 2> "$O/bench_$n.err"; python - <<PY
import json
d=json.load(open("$O/bench_$n.json"))
e=d["e2e"]; f=d.get("fabric",{})
print("$n", "e2e ms", round(e["ms_per_step"],3), "G", round(e["value"]/1e9,1), "transport", e["result_transport"], "other ms", round(e["other_result_transport"]["ms_per_step"],3), "fabric ms", round(f.get("ms_per_step",0),3), "frac", round(f.get("e2e_fraction_of_fabric",0),3), "u32 ms", round(d["e2e_from_cells"]["u32_as_is"]["ms_per_step"],2), "enc16 ms", round(d["e2e_from_cells"]["host_encode16"]["ms_per_step"],2), "dense", round(d.get("e2e_dense",{}).get("ms_per_step",0),2))
PY
}
run c4_default --config C4
run c4_slots8 --config C4 --slots 8
run c4_sites4 --config C4 --e2e-sites 37888
run c2_default
run c2_slots3 --slots 3
timeout 1200 python -m pytest tests -m gpu -q > "$O/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$O/pytest_gpu.log"; tail -3 "$O/pytest_gpu.log"
bash tools/gpu_cli.sh "$TAG" 2>&1 | tail -4 | cut -c1-900
