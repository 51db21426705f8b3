"""Warp instructions and stall samples of ONE kernel of an ncu report, grouped by code region: python tools/ncu_regions.py rep.ncu-rep kernel-substring
(the line ranges in bucket() follow bv_em_kernels.cuh as of the session that wrote profiles/r02_ncu_k4b_regions.txt; adjust them to the file)"""
import collections, csv, subprocess, sys
rep, want = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
kernel = hdr = cur = None
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
def bucket(f, l):
    if f == 'bv_em_kernels.cuh':
        if l < 310: return 'rcp/log_ratio helpers'
        if l < 335: return 'log_tab'
        if l < 372: return 'pick/put/sum_others'
        if l < 410: return 'em_bin'
        if l < 440: return 'em_sum_slow'
        if l < 500: return 'em_pass'
        if l < 545: return 'moved_bin/setup'
        if l < 600: return 'em_task loop'
        if l < 650: return 'em_task tail (lo loop, store)'
        if l < 735: return 'em_iter'
        if l < 870: return 'decide_site'
        return 'kernel body (staging, barriers)'
    if f in ('bv_fisher_fast.h', 'bv_math.cuh'): return 'fisher/gammaq/math'
    return f
for r in rows:
    if len(r) >= 2 and r[0] == "Function Name": kernel = r[1]
    elif len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No": hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-" and kernel and want in kernel:
        ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        k = bucket(cur, int(r[0]))
        agg[k][0] += int(r[ie]); agg[k][1] += int(r[isamp])
        for i, hn in enumerate(hdr):
            if hn.startswith("stall_") and "Not Issued" not in hn and r[i].isdigit() and int(r[i]): agg[k][2][hn[6:]] += int(r[i])
ti = sum(v[0] for v in agg.values()) or 1; ts = sum(v[1] for v in agg.values()) or 1
print(want, ti, ts)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:38s} inst {100*v[0]/ti:5.1f}% samp {100*v[1]/ts:5.1f}% {dict(v[2].most_common(4))}")
