"""Executed warp-instructions per code region of a kernel (regions delimited by BAR.SYNC / loop markers) from an
.ncu-rep source page: python tools/ncu_regions.py <rep> <kernel-regex> [n_units]   (n_units: e.g. matrices, to normalise)"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name"')
rows = list(csv.reader(io.StringIO('"Kernel Name"' + blocks[1])))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
seg, acc, n, total = 0, 0, 0, 0
marks = ("BAR.SYNC", "UTCBAR", "LDTM", "STTM", "EXIT")
for r in data:
    ins = int(r[ix['Instructions Executed']] or 0)
    acc += ins; n += 1; total += ins
    s = r[ix['Source']]
    if any(m in s for m in marks):
        print("region %2d  ends at %-40s static %4d  executed %12d  per-unit %8.1f" % (seg, s[:40], n, acc, acc / units))
        seg += 1; acc = 0; n = 0
print("tail static %d executed %d per-unit %.1f" % (n, acc, acc / units))
print("total executed %d per-unit %.1f" % (total, total / units))
