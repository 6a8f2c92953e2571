"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export by CUDA source line."""
import csv
import sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, out = None, None, []
for r in csv.reader(open(path)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[2] == "-":  # per-line aggregate rows have no SASS address
        try:
            out.append((cur, int(r[0]), r[1], int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")])))
        except ValueError:
            pass
ts, ti = sum(o[3] for o in out) or 1, sum(o[4] for o in out) or 1
print("total samples", ts, "warp instructions", ti)
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"{o[3]*100/ts:5.1f}% smp {o[4]*100/ti:5.1f}% ins | {o[0]}:{o[1]:<4d} {o[2].strip()[:120]}")
