"""Per-kernel key metrics (and the hottest source lines) of an ncu report that holds several kernels, e.g. one whole step:

    python tools/ncu_step_summary.py rep.ncu-rep [top_lines]   > profiles/<round>_ncu_full_<config>_step.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct"]
STALLS = "smsp__average_warps_issue_stalled_"

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
print(f"# {rep}: ncu --set full --clock-control none --import-source on (times under ncu are cold-cache and serialised)")
for r in rows[2:]:
    print(f"\n--- {r[ki]}")
    for i, h in enumerate(hdr):
        if h in KEYS:
            print(f"    {h:78s} {r[i]:>18s} {units[i]}")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(r[i]), h[len(STALLS):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    st.sort(reverse=True)
    print("    warps stalled per issue (top): " + ", ".join(f"{n} {v:.2f}" for v, n in st[:6]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
kernel = h2 = cur = None
agg = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0, ""]))
for r in csv.reader(src.splitlines()):
    if len(r) >= 2 and r[0] == "Function Name":
        kernel = r[1]
    elif len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No":
        h2 = r
    elif h2 and len(r) == len(h2) and r[0].isdigit() and r[2] == "-" and kernel:
        a = agg[kernel][(cur, int(r[0]))]
        a[0] += int(r[h2.index("Instructions Executed")])
        a[1] += int(r[h2.index("# Samples")])
        a[2] = r[1].strip()[:110]
for k, lines in agg.items():
    ti = sum(v[0] for v in lines.values()) or 1
    ts = sum(v[1] for v in lines.values()) or 1
    print(f"\n=== hottest source lines of {k[:60]} ({ti} warp instructions, {ts} stall samples)")
    for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"    {f}:{ln:<5d} inst {100 * v[0] / ti:5.1f} %  samples {100 * v[1] / ts:5.1f} %  | {v[2]}")
