set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
python bench.py --workload v1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_v1c_n1.json 2> gpurun_out/r02_bench_v1c_n1.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_v1c_n1.json').read().strip().splitlines()[-1]); r=d['roofline']; print('V1', round(d['value'],1), round(d['ms_per_step'],3), r['frac'], {k:v['ms'] for k,v in list(r['kernels'].items())[:8]})"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_v1_b16.csv python tools/profile_step.py 16 v1 > gpurun_out/ncu_v1.log 2>&1; tail -1 gpurun_out/ncu_v1.log | cut -c1-200
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm2_f16_kernel -c 2 -o gpurun_out/r02_v1_gemm_stage0 python tools/profile_step.py 16 v1 > gpurun_out/ncu_v1_gemm.log 2>&1; tail -1 gpurun_out/ncu_v1_gemm.log | cut -c1-200
