"""SASS-level summary of one kernel of an .ncu-rep (source page): instruction mix, stall reasons, hottest instructions.
    python tools/ncu_sass_summary.py <rep> <kernel-name-substring> [index]"""
import csv, io, subprocess, sys, collections, re
rep, pat = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name",')
sel = [b for b in blocks[1:] if pat in b.split("\n",1)[0]]
b = sel[which]
rows = list(csv.reader(io.StringIO(b.split("\n",1)[1])))
hdr = rows[0]; data = [r for r in rows[1:] if len(r) == len(hdr)]
ix = {h:i for i,h in enumerate(hdr)}
tot_i = sum(int(r[ix['Instructions Executed']]) for r in data)
tot_s = sum(int(r[ix['# Samples']]) for r in data)
print("kernel:", b.split("\n",1)[0][:90], " warp-instr", tot_i, " samples", tot_s)
byop = collections.Counter(); bys = collections.Counter()
for r in data:
    op = r[ix['Source']].strip().split()[0]
    if op.startswith('@'): op = r[ix['Source']].strip().split()[1]
    op = op.split('.')[0]
    byop[op] += int(r[ix['Instructions Executed']]); bys[op] += int(r[ix['# Samples']])
for op, c in byop.most_common(28):
    print("  %-12s instr %5.1f%%  samples %5.1f%%" % (op, 100*c/tot_i, 100*bys[op]/max(tot_s,1)))
stalls = [h for h in hdr if h.startswith('stall_')]
if stalls:
    tot = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
    print("  stalls:", ", ".join("%s %.1f%%" % (s[6:], 100*v/max(tot_s,1)) for s, v in sorted(tot.items(), key=lambda x:-x[1])[:8]))
print("  top sampled instructions:")
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:25]:
    print("   %6s %9s  %s   conflN=%s" % (r[ix['# Samples']], r[ix['Instructions Executed']], r[ix['Source']].strip()[:70], r[ix['L1 Conflicts Shared N-Way']]))
