"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.
usage: launch_summary.py launches.csv "<command that was profiled>" > summary.csv"""
import csv, sys, collections, re
path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = []
with open(path, newline="") as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.reader(lines)
hdr = None
for r in rd:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r
        continue
    if len(r) == len(hdr):
        rows.append(dict(zip(hdr, r)))
agg = collections.OrderedDict()
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
print("# ncu launch list of `%s`" % cmd)
print("# ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache/serialised: compare SHARES")
print("kernel,launches,total_ms,share")
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%s,%d,%.3f,%.1f%%" % (name, n, ms, 100. * ms / tot))
print("TOTAL,%d,%.3f,100%%" % (sum(a[0] for a in agg.values()), tot))
