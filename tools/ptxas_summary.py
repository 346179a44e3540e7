"""Summarise registers / spills / stack per kernel from the ptxas logs written by meta_learning_pacoh_b200/build.py."""
import glob, os, re, subprocess, sys
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "meta_learning_pacoh_b200", "build")
pat = sys.argv[1] if len(sys.argv) > 1 else ""
for f in sorted(glob.glob(os.path.join(root, "*.ptxas.log"))):
    blocks = re.split(r"ptxas info\s+: Compiling entry function ", open(f).read())[1:]
    for b in blocks:
        name = subprocess.run(["c++filt", b.split("'")[1]], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"pacoh::\(anonymous namespace\)::", "", name)[:60]
        if pat and not re.search(pat, name):
            continue
        regs = re.search(r"Used (\d+) registers", b).group(1)
        spill = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", b)
        stack = re.search(r"(\d+) bytes stack frame", b)
        print("%-62s regs=%-4s stack=%-4s spill=%s/%s" % (name, regs, stack.group(1) if stack else "-", spill.group(1), spill.group(2)))
