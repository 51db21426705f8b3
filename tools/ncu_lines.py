"""Summarise an ncu report by CUDA source line: instructions executed and stall samples.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file, hdr, items = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":
        items.append((cur_file, r))
ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
tot_i = sum(int(r[ie]) for _, r in items)
tot_s = sum(int(r[isamp]) for _, r in items)
print(f"total warp instructions {tot_i}, samples {tot_s}")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
for f, r in sorted(items, key=lambda x: -int(x[1][isamp]))[:top_n]:
    st = sorted(((int(r[i]), hdr[i][6:]) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:3]
    print(f"{f}:{r[0]:>4} inst {100 * int(r[ie]) / tot_i:5.1f}% samp {100 * int(r[isamp]) / tot_s:5.1f}%  "
          f"{' '.join(f'{n}:{c}' for c, n in st):40s} | {r[1].strip()[:90]}")
