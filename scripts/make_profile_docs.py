"""Turn ncu captures (gpurun_out/*.ncu-rep, launch-list csv) into the summaries kept under profiles/.

  python scripts/make_profile_docs.py full   <rep> <out.md> <title...>     # key metrics table of an `ncu --set full` capture
  python scripts/make_profile_docs.py launch <csv> <out.md>                # per-kernel totals of a gpu__time_duration launch list
  python scripts/make_profile_docs.py traffic <rep> <out.json>             # dram bytes read + written of the captured launch
"""
import collections, csv, json, subprocess, sys


def metrics(rep):
    out = subprocess.run([sys.executable, "scripts/ncu_summary.py", rep], capture_output=True, text=True).stdout
    rows = []
    for l in out.strip().splitlines():
        parts = l.split()
        name = parts[0]
        unit, val = (parts[1], parts[2]) if len(parts) >= 3 else ("", parts[1])
        try:
            float(unit); unit, val = "", unit
        except ValueError:
            pass
        rows.append((name, val, unit))
    return rows


def raw_bytes(rep, key):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines())); h = r[0]; i = h.index(key)
    return float(r[2][i]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[1][i], 1)


mode = sys.argv[1]
if mode == "full":
    rep, out, title = sys.argv[2], sys.argv[3], " ".join(sys.argv[4:])
    md = ["# " + title, "", "| metric | value | unit |", "|---|---|---|"] + ["| `%s` | %s | %s |" % r for r in metrics(rep)] + [""]
    open(out, "w").write("\n".join(md))
elif mode == "launch":
    rows = list(csv.reader(l for l in open(sys.argv[2]) if l.startswith('"')))
    hh = rows[0]; ki = hh.index("Kernel Name"); vi = hh.index("Metric Value")
    agg = collections.OrderedDict()
    for x in rows[1:]:
        a = agg.setdefault(x[ki].split("(")[0], [0, 0.0]); a[0] += 1; a[1] += float(x[vi].replace(",", "")) / 1e3
    tot = sum(a[1] for a in agg.values())
    md = ["| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    md += ["| %s | %d | %.1f | %.1f | %.3f |" % (n, a[0], a[1], a[1] / a[0], a[1] / tot) for n, a in sorted(agg.items(), key=lambda x: -x[1][1])]
    open(sys.argv[3], "w").write("\n".join(md) + "\n")
elif mode == "traffic":
    rd, wr = raw_bytes(sys.argv[2], "dram__bytes_read.sum"), raw_bytes(sys.argv[2], "dram__bytes_write.sum")
    json.dump({"dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes_per_launch": rd + wr}, open(sys.argv[3], "w"), indent=1)
