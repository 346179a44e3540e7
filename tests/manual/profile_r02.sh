# Round-2 profiles (run under gpurun, one GPU).  Nothing printed under ncu is a bench value.
set -x
export PACOH_GRAPH=0
# launch lists (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_c4.csv python bench.py --config 4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r02_c5.csv python bench.py --config 5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c5.log 2>&1
# full captures of the dominant kernels: config 4's GP kernel, config 5's tile kernels (one mid-size launch of each mode)
ncu --set full --clock-control none --import-source on -k regex:gp_tc_kernel -s 3 -c 1 -o gpurun_out/prof_r02_gptc python bench.py --config 4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gptc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:big_kernel -s 16 -c 2 -o gpurun_out/prof_r02_big_chol python tests/manual/big_profile.py 2048 296 > gpurun_out/ncu_big1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:big_kernel -s 131 -c 9 -o gpurun_out/prof_r02_big_inv python tests/manual/big_profile.py 2048 296 > gpurun_out/ncu_big2.log 2>&1
ncu --set full --clock-control none -k regex:"gp_post_kernel|pred_metrics" -c 4 -o gpurun_out/prof_r02_post python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ncu_post.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r02_*.csv
