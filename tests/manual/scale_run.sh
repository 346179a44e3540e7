# strong-scaling runs of bench.py on one box: bash tests/manual/scale_run.sh "8 4 2 1" [config] [extra args]
CFG=${2:-4}
for N in $1; do
  if [ "$N" = "1" ]; then CMD="python bench.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py"; fi
  timeout ${LIMIT:-240} $CMD --gpus $N --config $CFG --steps ${STEPS:-100} --warmup 10 --no-cpu-baseline $3 > gpurun_out/bench_r02_c${CFG}_n$N.json 2> gpurun_out/bench_r02_c${CFG}_n$N.err || tail -5 gpurun_out/bench_r02_c${CFG}_n$N.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_r02_c${CFG}_n$N.json'))
    print('N=$N config $CFG: %.4g evals/s  %.4f ms/step  e2e %.4g  %s | %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['cuda_graph'], d['config']['collective'][:60]))
    print('    ', {k:(round(v['ms_per_launch'],4)) for k,v in d['roofline']['kernels'].items()})
except Exception as e:
    print('N=$N failed', e)
PY
done
