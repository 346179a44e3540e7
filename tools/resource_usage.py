"""Per-kernel registers / shared memory / local-memory (spill) table of the built library, from `cuobjdump --dump-resource-usage`
(static: runs without a GPU).  Usage: python tools/resource_usage.py [path/to/libpacoh_b200.so] > profiles/resource_usage_rNN.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return [o if o else n for o, n in zip(out, names)]


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "meta_learning_pacoh_b200", "libpacoh_b200.so")
    txt = subprocess.run(["cuobjdump", "--dump-resource-usage", so], capture_output=True, text=True, check=True).stdout
    rows = []
    for m in re.finditer(r"Function (\S+):\n\s*(.*)", txt):
        res = dict(kv.split(":") for kv in m.group(2).split() if ":" in kv)
        rows.append((m.group(1), int(res.get("REG", 0)), int(res.get("SHARED", 0)), int(res.get("LOCAL", 0)), int(res.get("STACK", 0))))
    names = demangle([r[0] for r in rows])
    short = [re.sub(r"\(.*", "", n.replace("(anonymous namespace)::", "")).replace("void ", "") for n in names]
    print("# %s: %d kernels, sm_100a (cuobjdump --dump-resource-usage; SHARED is the static part, dynamic shared memory is set at launch)"
          % (os.path.basename(so), len(rows)))
    print("%-92s %5s %8s %6s %6s" % ("kernel", "REG", "SHARED", "LOCAL", "STACK"))
    for s, r in sorted(zip(short, rows)):
        print("%-92s %5d %8d %6d %6d" % (s[:92], r[1], r[2], r[3], r[4]))
    spilled = sorted({"%s (%d B)" % (s, r[3] + r[4]) for s, r in zip(short, rows) if r[3] + r[4] > 0})
    print("# kernels with a stack frame / local memory (register spills under __launch_bounds__, or indexed local arrays): %d of %d"
          % (len(spilled), len(rows)))
    for s in spilled:
        print("#   " + s)

if __name__ == "__main__":
    main()
