set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "PARITY|passed|failed|FAILED" > gpurun_out/r02_parity_gpu.log; tail -12 gpurun_out/r02_parity_gpu.log | grep -E "passed|failed|FAILED"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err; tail -c 300 gpurun_out/r02_bench_final_n1.json
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_final_n1_b.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_final_n1_b.json').read().strip().splitlines()[-1]); r=d['roofline']; print('default', round(d['value'],1), round(d['ms_per_step'],3), r['frac'], r['kernels']['gemm_f16_kernel']['ms'], r['kernels']['attn_fwd_kernel']['ms'])"
python bench.py --workload hires --steps 10 --warmup 3 > gpurun_out/r02_bench_final_hires.json 2> gpurun_out/r02_bench_final_hires.err; tail -c 200 gpurun_out/r02_bench_final_hires.json
python bench.py --workload v1 --steps 10 --warmup 3 > gpurun_out/r02_bench_final_v1.json 2> gpurun_out/r02_bench_final_v1.err; tail -c 200 gpurun_out/r02_bench_final_v1.json
UDB_LIB=/root/repo/variants/libudb_trace.so timeout 300 python tools/attn_trace.py 2>&1 | tail -12 | head -6 | tee gpurun_out/r02_attn2_trace.txt
timeout 600 python tools/bench_kernels.py attn 2>&1 | grep -v Warn | tee gpurun_out/r02_kernels_attn.txt
timeout 600 python tools/bench_kernels.py gemm 2>&1 | grep -v Warn | tee gpurun_out/r02_kernels_gemm.txt
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_step_b8.csv python tools/profile_step.py 8 default > /dev/null 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_fwd -c 1 -o gpurun_out/r02_attn python tools/profile_step.py 8 default > /dev/null 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_fwd -c 1 -o gpurun_out/r02_attn_hires python tools/profile_step.py 4 hires > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
