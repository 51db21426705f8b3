"""Aggregate an ncu source page by (file, function-ish region): python tools/ncu_funcs.py rep.ncu-rep
Regions are found by scanning the CUDA source for `__device__`/`__global__` function starts."""
import csv, re, subprocess, sys, os, collections
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur, hdr, items = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1]
    elif len(r) > 8 and r[0] == "Line No": hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-": items.append((cur, r))
ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
tot_i = sum(int(r[ie]) for _, r in items); tot_s = sum(int(r[isamp]) for _, r in items)
starts = {}
def regions(path):
    if path in starts: return starts[path]
    reg = []
    try:
        for n, ln in enumerate(open(path), 1):
            m = re.match(r"\s*(?:static\s+)?(?:__device__|__global__).*?(\w+)\s*\(", ln)
            if m and not ln.strip().startswith("//"): reg.append((n, m.group(1)))
    except OSError: pass
    starts[path] = reg; return reg
agg = collections.defaultdict(lambda: [0, 0])
for f, r in items:
    local = f
    if not os.path.exists(local):
        k = f.find("basevar_b200/"); local = f[k:] if k >= 0 else f
    name = "?"
    for n, nm in regions(local):
        if n <= int(r[0]): name = nm
    key = f"{os.path.basename(f)}:{name}"
    agg[key][0] += int(r[ie]); agg[key][1] += int(r[isamp])
print(f"total warp instructions {tot_i}, samples {tot_s}")
for k, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:50s} inst {100*i/tot_i:5.1f}%  samp {100*s/tot_s:5.1f}%")
