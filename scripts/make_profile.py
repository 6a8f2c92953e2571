"""profiles/rN from ncu reports: per-kernel summary csv, traffic.json entries, per-source-line shares, SASS opcode histogram.

    python scripts/make_profile.py profiles/r2 step=gpurun_out/r2_step_full.ncu-rep pile=gpurun_out/r2_pile_tile_d.ncu-rep
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    out = sys.argv[1]
    os.makedirs(out, exist_ok=True)
    traffic = {}
    for arg in sys.argv[2:]:
        tag, rep = arg.split("=")
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        h, units = rows[0], rows[1]
        with open(os.path.join(out, f"ncu_full_{tag}_summary.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["kernel"] + KEYS)
            w.writerow(["unit"] + [units[h.index(k)] if k in h else "" for k in KEYS])
            for r in rows[2:]:
                name = r[h.index("Kernel Name")]
                w.writerow([name] + [r[h.index(k)] if k in h else "" for k in KEYS])
                short = re.sub(r"^void ", "", name).split("(")[0]
                traffic.setdefault(tag, {})[short] = {
                    "dram_read_bytes": to_bytes(r[h.index("dram__bytes_read.sum")], units[h.index("dram__bytes_read.sum")]),
                    "dram_write_bytes": to_bytes(r[h.index("dram__bytes_write.sum")], units[h.index("dram__bytes_write.sum")]),
                    "duration_us": float(r[h.index("gpu__time_duration.sum")]),
                    "issue_active_pct": float(r[h.index("smsp__issue_active.avg.pct_of_peak_sustained_active")]),
                    "fp64_pipe_pct": float(r[h.index("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")]),
                    "alu_pipe_pct": float(r[h.index("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")]),
                    "warp_instructions": float(r[h.index("smsp__inst_executed.sum")]), "registers": int(float(r[h.index("launch__registers_per_thread")]))}
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
        tmp = os.path.join(out, f".{tag}_src.csv")
        open(tmp, "w").write(src)
        lines = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_lines.py"), tmp, "60"], capture_output=True, text=True).stdout
        open(os.path.join(out, f"{tag}_source_lines.txt"), "w").write(lines)
        os.remove(tmp)
    json.dump(traffic, open(os.path.join(out, "ncu_counters.json"), "w"), indent=1)
    # SASS opcode histogram of every kernel in the library
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "crisp_b200", "csrc", "libcrispgpu.so")], capture_output=True, text=True).stdout
    hist, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
            hist[cur] = {}
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            hist[cur][m.group(1)] = hist[cur].get(m.group(1), 0) + 1
    with open(os.path.join(out, "sass_opcodes.txt"), "w") as f:
        f.write("SASS opcode histogram per kernel (cuobjdump -sass crisp_b200/csrc/libcrispgpu.so, sm_100a); static instruction counts\n")
        for k in sorted(hist):
            tot = sum(hist[k].values())
            top = sorted(hist[k].items(), key=lambda kv: -kv[1])
            f.write(f"\n{k}: {tot} instructions\n  " + "  ".join(f"{op} {n}" for op, n in top[:28]) + "\n")
            special = {op: n for op, n in hist[k].items() if op in ("DMMA", "DFMA", "DADD", "DMUL", "MUFU", "ATOMS", "ATOMG", "RED", "REDUX", "SHFL", "LDGSTS", "UTMALDG", "IMAD", "LOP3", "PRMT", "POPC", "FLO")}
            f.write("  of note: " + ", ".join(f"{op} {n}" for op, n in sorted(special.items())) + "\n")


if __name__ == "__main__":
    main()
