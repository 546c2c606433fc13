"""Top source lines of an .ncu-rep source page dump (ncu --page source --csv --print-source cuda,sass)."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur = None; hdr = None; out = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[0] == "": continue
    d = dict(zip(hdr[4:], r[4:]))
    try:
        out.append((int(d["# Samples"]), int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), cur, r[0], r[1].strip()[:110]))
    except ValueError:
        pass
ts = sum(o[0] for o in out); ti = sum(o[1] for o in out)
print("total samples %d, warp insts %d" % (ts, ti))
for o in sorted(out, reverse=True)[:top]:
    print("%5.2f%% smp %5.2f%% inst  thr/inst %4.1f  %s:%s  %s" % (100.0 * o[0] / ts, 100.0 * o[1] / ti, o[2] / max(1, o[1]), o[3], o[4], o[5]))
