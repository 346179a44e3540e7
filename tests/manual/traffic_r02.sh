export PACOH_GRAPH=0
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/traffic_c5.csv python bench.py --config 5 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/traffic_c4.csv python bench.py --config 4 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv, collections, json
out = {}
for cfg, first in ((5, 'mlp_tc_fwd'), (4, 'step_prepare')):
    rows = list(csv.reader(l for l in open('gpurun_out/traffic_c%d.csv' % cfg) if l.startswith('"')))
    hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
    recs = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) != len(hdr): continue
        rid = int(r[ix['ID']])
        d = recs.setdefault(rid, {'name': r[ix['Kernel Name']]})
        v = float(r[ix['Metric Value']]); u = r[ix['Metric Unit']]
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0}.get(u, 1)
        d[r[ix['Metric Name']]] = v * scale
    ids = sorted(recs)
    starts = [i for i in ids if first in recs[i]['name']]
    s0, s1 = starts[-2], starts[-1]
    stage = collections.OrderedDict()
    for i in ids:
        if not (s0 <= i < s1): continue
        n = recs[i]['name']
        key = 'mlp_fwd' if 'mlp_tc_fwd' in n else 'mlp_bwd' if 'mlp_tc_bwd' in n else 'gp_mll' if ('big_' in n or 'gp_tc' in n or 'gp_mll' in n) else 'other'
        st = stage.setdefault(key, {'launches': 0, 'dram_bytes': 0.0, 'ms': 0.0})
        st['launches'] += 1
        st['dram_bytes'] += recs[i].get('dram__bytes_read.sum', 0) + recs[i].get('dram__bytes_write.sum', 0)
        st['ms'] += recs[i].get('gpu__time_duration.sum', 0)
    out['config%d' % cfg] = {k: v['dram_bytes'] for k, v in stage.items()}
    out['config%d_detail' % cfg] = stage
    print(cfg, json.dumps(stage))
json.dump(out, open('gpurun_out/traffic_r02.json', 'w'), indent=1)
PY
