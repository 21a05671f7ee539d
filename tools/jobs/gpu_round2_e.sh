set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_c_n1.json 2> gpurun_out/r02_bench_c_n1.err; tail -c 1800 gpurun_out/r02_bench_c_n1.json; tail -3 gpurun_out/r02_bench_c_n1.err
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_v1_b16.csv python tools/profile_step.py 16 v1 > gpurun_out/ncu_v1.log 2>&1; tail -2 gpurun_out/ncu_v1.log | cut -c1-300
