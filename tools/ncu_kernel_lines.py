"""Hot source lines of ONE kernel in an ncu report: python tools/ncu_kernel_lines.py rep.ncu-rep kernel-substring [top_n]"""
import collections
import csv
import subprocess
import sys

rep, want = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
kernel = hdr = cur = None
agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), ""])
for r in rows:
    if len(r) >= 2 and r[0] == "Function Name":
        kernel = r[1]
    elif len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-" and kernel and want in kernel:
        ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        k = (cur, int(r[0]))
        agg[k][0] += int(r[ie]); agg[k][1] += int(r[isamp]); agg[k][3] = r[1][:100]
        for i, hn in enumerate(hdr):
            if hn.startswith("stall_") and "Not Issued" not in hn and r[i].isdigit() and int(r[i]):
                agg[k][2][hn[6:]] += int(r[i])
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
print(f"{want}: warp instructions {ti}, samples {ts}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1 if "--by-samples" in sys.argv else 0])[:top_n]:
    print(f"{k[0]}:{k[1]:4d} inst {100 * v[0] / ti:5.1f}% samp {100 * v[1] / ts:5.1f}% {dict(v[2].most_common(2))} | {v[3]}")
