# usage: bash tests/manual/bench_all.sh "2 3 5" [extra bench args]   (writes gpurun_out/bench_r02_c<N>.json)
for c in $1; do timeout 900 python bench.py --config $c --steps ${STEPS:-40} --warmup 5 $2 > gpurun_out/bench_r02_c$c.json 2> gpurun_out/bench_r02_c$c.err || tail -5 gpurun_out/bench_r02_c$c.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_r02_c$c.json'))
    print('config $c: value %.4g evals/s  ms/step %.4f  e2e %.4g (%.4f ms)  graph=%s  launches/step %d  dom=%s frac=%.3f step_frac=%.3f cpu=%s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['cuda_graph'], d['gpu_launches']/d['steps'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['step_frac_of_fp32_peak'], d['cpu_baseline'] and (d['cpu_baseline']['value'], d['cpu_baseline']['single_thread']['value'])))
    print('   ', {k:(round(v['ms_per_launch'],4)) for k,v in d['roofline']['kernels'].items()}, d['clocks'])
except Exception as e:
    print('config $c failed', e)
PY
done
