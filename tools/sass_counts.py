"""Per-kernel SASS instruction counts of the shipped library (evidence for DESIGN.md: which hardware paths a kernel uses).

    python tools/sass_counts.py [basevar_b200/libbasevar_b200.so] > profiles/r02_sass_counts.txt

UBLKCP = TMA bulk copy (cp.async.bulk), SYNCS = mbarrier, IDP.4A = dp4a, ATOMS / REDS = shared-memory atomics, D* = FP64 pipe,
MUFU.RCP64H = the reciprocal estimate of rcp_fast / of divisions.  No UTMALDG / UTC*MMA / HMMA is expected: the path is a
byte-streaming reduction, not a contraction.
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "basevar_b200/libbasevar_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
KEYS = ["UBLKCP", "SYNCS", "IDP.4A", "ATOMS", "REDS", "ATOMG", "REDG", "RED.", "DFMA", "DADD", "DMUL", "DSETP", "MUFU.RCP64H", "MUFU", "LDS", "STS", "LDG", "STG",
        "LDL", "STL", "SHFL", "VOTE", "REDUX", "BAR", "CALL", "UTMALDG", "UTCHMMA", "HMMA"]
cnt = collections.OrderedDict()
name = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        cnt[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        cnt[name]["total"] += 1
        for k in KEYS:
            if op.startswith(k) or (k.endswith(".") and op.startswith(k[:-1] + ".")):
                cnt[name][k] += 1
print("library:", lib, "| architectures in the fatbin:", ", ".join(arch))
cols = ["total"] + [k for k in KEYS if any(c[k] for c in cnt.values())]
print("%-46s" % "kernel" + "".join("%9s" % c[:9] for c in cols))
for n, c in cnt.items():
    print("%-46s" % n[-46:] + "".join("%9d" % c[k] for k in cols))
