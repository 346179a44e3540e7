set -x
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_r02e.log 2>&1; grep -E "^E +|^FAILED|passed|failed" gpurun_out/pytest_r02e.log | head -30
for c in 2 3 1 4; do timeout 600 python bench.py --config $c --steps 40 --warmup 5 > gpurun_out/bench_r02_c$c.json 2> gpurun_out/bench_r02_c$c.err; tail -3 gpurun_out/bench_r02_c$c.err; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_r02_c$c.json'))
print('config $c: value %.4g evals/s  ms/step %.4f  e2e %.4g  graph=%s  launches/step %d  dom=%s frac=%.3f  cpu=%s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['cuda_graph'], d['gpu_launches']/d['steps'], d['roofline']['kernel'], d['roofline']['frac'], d['cpu_baseline'] and d['cpu_baseline']['value']))
print({k:(round(v['ms_per_launch'],4)) for k,v in d['roofline']['kernels'].items()})
"; done
