"""Writes profiles/asm_traffic.json from an `ncu --set full` capture of the assembly kernel:
    python scripts/ncu_traffic.py profiles/r2_k_numeric_brick3_256cubed.ncu-rep 256
bench.py reports these DRAM bytes as roofline.traffic while the kernel sources still have the recorded MD5."""
import csv, hashlib, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, n = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
row = [r for r in data if "k_numeric_brick3" in r[hdr.index("Kernel Name")]][0]
def val(name):
    i = hdr.index(name)
    v, u = float(row[i]), units[i]
    return v * {"byte": 1., "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
sources = ["mimpy_b200/csrc/assemble_brick3.cu", "mimpy_b200/csrc/cell_core.cuh", "mimpy_b200/csrc/assemble_args.cuh",
           "mimpy_b200/csrc/common.cuh"]
h = hashlib.md5()
for f in sources:
    h.update(open(os.path.join(ROOT, f), "rb").read())
rec = {"kernel": row[hdr.index("Kernel Name")], "n": n, "dram_bytes_read": val("dram__bytes_read.sum"),
       "dram_bytes_write": val("dram__bytes_write.sum"), "gpu_time_ms": float(row[hdr.index("gpu__time_duration.sum")]),
       "capture": os.path.relpath(rep, ROOT), "sources": sources, "sources_md5": h.hexdigest()}
json.dump(rec, open(os.path.join(ROOT, "profiles", "asm_traffic.json"), "w"), indent=1)
print(rec)
