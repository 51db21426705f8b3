"""Aggregate an ncu source page by function region, per kernel: python tools/ncu_funcs.py rep.ncu-rep [kernel-substring]"""
import csv, re, subprocess, sys, os, collections
rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
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
kernel, cur, hdr = None, None, None
per_kernel = collections.OrderedDict()
for r in rows:
    if len(r) >= 2 and r[0] == "Function Name": kernel = r[1].split("(")[0]; per_kernel.setdefault(kernel, [])
    elif len(r) >= 2 and r[0] == "File Path": cur = r[1]
    elif len(r) > 8 and r[0] == "Line No": hdr = r; ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    elif hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":
        rr = list(r); rr[7], rr[6] = r[ie], r[isamp]
        per_kernel.setdefault(kernel, []).append((cur, rr))
ie, isamp = 7, 6
for kernel, items in per_kernel.items():
    if want not in (kernel or ""): continue
    tot_i = sum(int(r[ie]) for _, r in items) or 1; tot_s = sum(int(r[isamp]) for _, r in items) or 1
    agg = collections.defaultdict(lambda: [0, 0])
    for f, r in items:
        local = f
        if not os.path.exists(local):
            k = f.find("basevar_b200/"); local = f[k:] if k >= 0 else f
        name = "?"
        for n, nm in regions(local):
            if n <= int(r[0]): name = nm
        agg[f"{os.path.basename(f)}:{name}"][0] += int(r[ie]); agg[f"{os.path.basename(f)}:{name}"][1] += int(r[isamp])
    print(f"== {kernel}: warp instructions {tot_i}, samples {tot_s}")
    for k, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:24]:
        print(f"   {k:48s} inst {100*i/tot_i:5.1f}%  samp {100*s/tot_s:5.1f}%")
