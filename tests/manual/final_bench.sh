set -x
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_r02_final.log 2>&1; tail -3 gpurun_out/pytest_r02_final.log
STEPS=3000 bash tests/manual/bench_all.sh "2 3 1"
STEPS=100 bash tests/manual/bench_all.sh "4"
for n in 512 1024 2048; do
  timeout 900 python bench.py --config 5 --points $n --steps 20 --warmup 5 > gpurun_out/bench_r02_c5_$n.json 2> gpurun_out/bench_r02_c5_$n.err || tail -5 gpurun_out/bench_r02_c5_$n.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_r02_c5_$n.json') if l.startswith('{')][-1])
print('config 5 n=$n: value %.4g evals/s  ms/step %.3f  e2e %.4g  dom=%s frac=%.3f  step_frac=%.3f  cpu=%s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['step_frac_of_fp32_peak'], (d['cpu_baseline']['value'], d['cpu_baseline']['single_thread']['value'])))
print('   ', {k:(round(v['ms_per_launch'],4)) for k,v in d['roofline']['kernels'].items()}, d['clocks'])
PY
done
