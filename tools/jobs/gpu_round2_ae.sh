set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_last_n1.json 2> gpurun_out/r02_bench_last_n1.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_last_n1.json').read().strip().splitlines()[-1]); r=d['roofline']; print('default', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), r['frac'], r['kernels']['gemm_f16_kernel']['ms'], r['kernels']['attn_fwd_kernel']['ms'], r['kernels']['small_linear_kernel']['ms'], d['clocks'])"
