"""Census of the Blackwell / Hopper data-movement instructions in the shipped library, per kernel:
UBLKCP (cp.async.bulk), SYNCS (mbarrier), USETMAXREG (setmaxnreg), UTMALDG / UTMASTG (tensor-map TMA), and the
Ampere-style LDGSTS (cp.async).  usage: python scripts/sass_census.py [library] -> text on stdout"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ("UBLKCP", "SYNCS", "USETMAXREG", "UTMALDG", "UTMASTG", "LDGSTS")


def census(lib=None):
    lib = lib or os.path.join(ROOT, "mimpy_b200", "libmimpy_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    current, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            current = m.group(1)
            continue
        for op in OPS:
            if op in line and current is not None:
                counts.setdefault(current, collections.Counter())[op] += 1
    return counts


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout
        return out.splitlines()
    except Exception:
        return list(names)


if __name__ == "__main__":
    c = census(sys.argv[1] if len(sys.argv) > 1 else None)
    total = collections.Counter()
    for name, pretty in zip(c, demangle(list(c))):
        total.update(c[name])
        print("%-110s %s" % (pretty[:110], "  ".join("%s %d" % (op, c[name][op]) for op in OPS if c[name][op])))
    print("whole library: " + "  ".join("%s %d" % (op, total[op]) for op in OPS))
