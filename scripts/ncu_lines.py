"""Per-source-line table of an ncu source page (ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > f.csv).

usage: python scripts/ncu_lines.py f.csv [min_pct]
Prints, per file, every line with at least min_pct % of the warp instructions: samples %, warp-instruction %,
thread-instruction %, lanes per instruction — the view that shows which part of a kernel runs with how many lanes.
"""
import csv
import sys

path = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
cur = None
hdr = None
rows = []
for r in csv.reader(open(path)):
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
        hdr = None
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    d = {}
    for k, v in zip(hdr[4:], r[4:]):
        d[k] = v
    try:
        rows.append((cur, int(r[0]), r[1].strip(), int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0), int(d["Thread Instructions Executed"] or 0)))
    except ValueError:
        pass
ts = sum(x[3] for x in rows) or 1
ti = sum(x[4] for x in rows) or 1
tt = sum(x[5] for x in rows) or 1
print("total: samples %d, warp instructions %.3f G, thread instructions %.3f G, lanes/inst %.2f" % (ts, ti / 1e9, tt / 1e9, tt / ti))
for f in sorted(set(x[0] for x in rows)):
    fi = sum(x[4] for x in rows if x[0] == f)
    ft = sum(x[5] for x in rows if x[0] == f)
    print("== %s: %.1f %% of warp instructions, %.1f %% of thread instructions, %.1f lanes" % (f, 100.0 * fi / ti, 100.0 * ft / tt, ft / max(1, fi)))
    for x in rows:
        if x[0] == f and 100.0 * x[4] / ti >= min_pct:
            print("%5d  smp %5.2f%%  winst %5.2f%%  tinst %5.2f%%  lanes %4.1f  %s" % (x[1], 100.0 * x[3] / ts, 100.0 * x[4] / ti, 100.0 * x[5] / tt, x[5] / max(1, x[4]), x[2][:120]))
