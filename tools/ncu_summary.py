"""Summarise an .ncu-rep (raw page + source page) into text: python tools/ncu_summary.py <rep> [kernel-regex]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.max']
for r in data:
    name = r[ix['Kernel Name']]
    print("== kernel:", name[:100])
    for w in want:
        if w in ix:
            print("   %-70s %s %s" % (w, r[ix[w]], units[ix[w]]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-name", "regex:" + kre] if kre else []),
                     capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name"')
for b in blocks[1:]:
    rows = list(csv.reader(io.StringIO('"Kernel Name"' + b)))
    print("== source-level stalls:", rows[0][1][:90])
    hdr, data = rows[1], [r for r in rows[2:] if len(r) == len(rows[1])]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    total = sum(int(r[ix['# Samples']] or 0) for r in data)
    ins = sum(int(r[ix['Instructions Executed']] or 0) for r in data)
    print("   warp instructions executed: %d   samples: %d" % (ins, total))
    tot = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
    for s, v in sorted(tot.items(), key=lambda x: -x[1])[:8]:
        print("   %-26s %6.2f %%" % (s, 100.0 * v / max(total, 1)))
    top = sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:14]
    for r in top:
        dom = max(stalls, key=lambda s: int(r[ix[s]] or 0))
        print("   %6s  %-58s %s" % (r[ix['# Samples']], r[ix['Source']].strip()[:58], dom))
