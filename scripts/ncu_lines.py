"""Per-CUDA-source-line instruction and stall-sample shares of one kernel of an .ncu-rep.
usage: ncu_lines.py rep.ncu-rep file.cubin mangled_kernel_name [top]
(ncu's SASS page is joined with `nvdisasm -g` line info by instruction offset)"""
import csv, io, re, subprocess, sys, collections
rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if "Address" in r and "# Samples" in r][0]
h = rows[hi]
ia, ii, isamp = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
sass = []
for r in rows[hi + 1:]:
    try:
        sass.append((int(r[ia], 16), int(r[ii]), int(r[isamp]), r[h.index("Source")]))
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
agg = collections.defaultdict(lambda: [0, 0])
ti = ts = 0
for a, n, s, src in sass:
    k = off2line.get(a - base)
    agg[k][0] += n; agg[k][1] += s; ti += n; ts += s
print("total warp instructions %d, samples %d" % (ti, ts))
for k, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% samples  %s" % (100. * n / ti, 100. * s / max(ts, 1), k))
