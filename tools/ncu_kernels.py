"""Key metrics of every kernel launch in an ncu report: python tools/ncu_kernels.py rep.ncu-rep"""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__pcsamp_warps_issue_stalled_no_instructions', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active', 'sass__inst_executed_local_loads', 'lts__t_sector_hit_rate.pct']
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h, u = rows[0], rows[1]
for r in rows[2:]:
    print("---", r[h.index("Kernel Name")].split("(")[0])
    for k in KEYS:
        if k in h:
            print(f"    {k:66s} {r[h.index(k)][:24]:>24s} {u[h.index(k)]}")
