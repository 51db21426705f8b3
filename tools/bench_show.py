"""Digest of bench.py JSON lines: python tools/bench_show.py file.json ..."""
import json
import sys

for fn in sys.argv[1:]:
    try:
        d = json.loads(open(fn).read().strip().splitlines()[-1])
    except Exception as e:
        print(fn, "no bench line", e)
        continue
    print("==", fn)
    print("value %.4g ms %.4f frac %.3f | %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["config"]["workload"][:60]))
    print("kernel_ms", {k: round(v, 4) for k, v in d["roofline"]["kernel_ms"].items()})
    for k in ("e2e", "e2e_from_cells", "e2e_dense", "fabric", "cpu_baseline", "cli"):
        if k in d:
            print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in d[k].items() if not isinstance(b, (str, dict)) or (isinstance(b, str) and len(b) < 40)})
            for a, b in d[k].items():
                if isinstance(b, dict):
                    print("   ", a, {x: (round(y, 4) if isinstance(y, float) else y) for x, y in b.items() if not isinstance(y, str) or len(y) < 40})
    for k, c in d.get("configs", {}).items():
        print(k, "value %.4g ms %.4f frac %.3f" % (c["value"], c["ms_per_step"], c["roofline"]["frac"]),
              {a: round(b, 4) for a, b in c["roofline"]["kernel_ms"].items()}, "cpu", c.get("cpu_baseline", {}).get("value"))
