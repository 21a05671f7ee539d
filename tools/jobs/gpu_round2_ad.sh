set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_small.py 2>&1 | grep -v "^$" | tail -25) | tee gpurun_out/r02_sanitizer_memcheck.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_final_n1.json').read().strip().splitlines()[-1]); r=d['roofline']; print('default', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), r['frac'], r['kernels']['gemm_f16_kernel']['ms'], r['kernels']['attn_fwd_kernel']['ms'], r['kernels']['small_linear_kernel'])"
python bench.py --workload hires --steps 10 --warmup 3 > gpurun_out/r02_bench_final_hires.json 2>/dev/null; tail -c 150 gpurun_out/r02_bench_final_hires.json
python bench.py --workload v1 --steps 10 --warmup 3 > gpurun_out/r02_bench_final_v1.json 2>/dev/null; tail -c 150 gpurun_out/r02_bench_final_v1.json
