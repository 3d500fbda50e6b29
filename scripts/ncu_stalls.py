"""Per-source-line stall-reason breakdown of one kernel of an .ncu-rep.
usage: ncu_stalls.py rep.ncu-rep file.cubin mangled_kernel_name [top]"""
import csv, io, re, subprocess, sys, collections
rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if "Address" in r and "# Samples" in r][0]
h = rows[hi]
reasons = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
ia, ii = h.index("Address"), h.index("Instructions Executed")
sass = []
for r in rows[hi + 1:]:
    try:
        sass.append((int(r[ia], 16), int(r[ii]), {c: int(r[h.index(c)] or 0) for c in reasons}, r[h.index("Source")]))
    except Exception:
        pass
base = sass[0][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
inside, line, off2line = False, None, {}
for l in dis:
    if l.startswith("\t.section") or l.startswith("//----"):
        inside = ("text." + kname) in l if "text." in l else inside
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/", l)
    if m:
        off2line[int(m.group(1), 16)] = line
tot = collections.Counter()
per = collections.defaultdict(collections.Counter)
for a, n, st, src in sass:
    k = off2line.get(a - base)
    for c, v in st.items():
        tot[c] += v
        per[k][c] += v
allsum = sum(tot.values())
print("stall samples by reason:", ", ".join("%s %.1f%%" % (c[6:], 100. * v / allsum) for c, v in tot.most_common(9)))
for k, cnt in sorted(per.items(), key=lambda kv: -sum(kv[1].values()))[:top]:
    s = sum(cnt.values())
    print("%5.1f%%  %-32s %s" % (100. * s / allsum, k, ", ".join("%s %.1f" % (c[6:], 100. * v / allsum) for c, v in cnt.most_common(4))))
