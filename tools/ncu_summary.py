"""Key metrics of an ncu report (one kernel launch): python tools/ncu_summary.py rep.ncu-rep"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size",
    "sm__inst_executed_pipe_uniform.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_cbu.sum.pct_of_peak_sustained_active",
]
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
for i, h in enumerate(hdr):
    if h in KEYS or ("warp_issue_stalled" in h and h.endswith("_per_warp_active.pct")):
        try:
            if "stalled" in h and float(vals[i]) < 3:
                continue
        except ValueError:
            pass
        print(f"{h:90s} {vals[i]:>16s} {units[i]}")
