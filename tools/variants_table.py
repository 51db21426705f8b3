"""Table of a tools/gpu_parity_timings.sh / gpu_variants.sh log: python tools/variants_table.py gpurun_out/<tag>/run_kernel.log"""
import re
import sys

cfg = var = None
for l in open(sys.argv[1]):
    l = l.strip()
    if l.startswith("default:"):
        cfg, var = l[9:], "default"
        continue
    if l.startswith("variant"):
        var = l.split()[1].rstrip(":")
        continue
    m = re.search(r"launches ms \[(.*?)\].*kernel ms (.*?); bound.*K4 = (.*?); flags", l)
    if m:
        ms = [float(x.strip("' ")) for x in m.group(1).split(",")]
        print(f"{cfg[9:40]:32s} {var:8s} best {min(ms[1:]):.3f}  {m.group(2)}  | {m.group(3)}")
