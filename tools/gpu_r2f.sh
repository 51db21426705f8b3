#!/usr/bin/env bash
set -u
TAG="${1:-r2f}"; VARS="${2:-}"
O=gpurun_out/$TAG; mkdir -p "$O"
bash tools/gpu_r2a.sh "$TAG" "$VARS"
timeout 900 python bench.py > "$O/bench.json" 2> "$O/bench.err"; echo "bench rc=$?"; tail -c 1500 "$O/bench.err"
python - <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1] if len(sys.argv)>1 else "gpurun_out/r2f/bench.json").read().strip().splitlines()[-1])
except Exception as e:
    print("no bench line", e); sys.exit(0)
print("value %.4g ms %.4f frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"]))
print("kernel_ms", {k: round(v,4) for k,v in d["roofline"]["kernel_ms"].items()})
for k in ("e2e","e2e_from_cells","e2e_sparse_u32","e2e_dense","fabric","cpu_baseline"):
    if k in d: print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in d[k].items() if not isinstance(b,str) or len(b)<40})
for k,c in d.get("configs",{}).items():
    print(k, "value %.4g ms %.4f frac %.3f" % (c["value"], c["ms_per_step"], c["roofline"]["frac"]), {a: round(b,4) for a,b in c["roofline"]["kernel_ms"].items()}, "cpu", c.get("cpu_baseline",{}).get("value"))
PY
