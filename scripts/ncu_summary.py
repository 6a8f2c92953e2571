"""Key counters of every kernel in an ncu report (csv of `ncu -i X.ncu-rep --page raw --csv` on stdin)."""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
rows = list(csv.reader(sys.stdin))
h = rows[0]
for r in rows[2:]:
    print("==", r[h.index("Kernel Name")][:110])
    for k in KEYS:
        if k in h:
            print(f"  {k:72s} {r[h.index(k)]:>16s} {rows[1][h.index(k)]}")
    st = []
    for name in ("long_scoreboard", "short_scoreboard", "barrier", "wait", "not_selected", "math_pipe_throttle", "mio_throttle", "lg_throttle",
                 "branch_resolving", "no_instruction", "dispatch_stall", "drain", "imc_miss", "tex_throttle", "sleeping", "membar"):
        k = STALLS % name
        if k in h:
            st.append((float(r[h.index(k)]), name))
    print("  stalls per issue:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:7]))
