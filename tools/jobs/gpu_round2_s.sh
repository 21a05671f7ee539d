set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "PARITY|passed|failed|FAILED" > gpurun_out/r02_parity_gpu_poly.log; tail -8 gpurun_out/r02_parity_gpu_poly.log
b() { tag=$1; shift; env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_s_$tag.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_s_$tag.json').read().strip().splitlines()[-1]); r=d['roofline']; print('$tag', round(d['value'],1), round(d['ms_per_step'],3), r['frac'], r['kernels']['gemm_f16_kernel']['ms'], r['kernels']['attn_fwd_kernel']['ms'])"; }
for i in 1 2 3; do
b bnauto_$i X=1
b bn256_$i UDB_GEMM_BN_AUTO=0
done
b v1_auto X=1 
python bench.py --workload v1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('v1 auto', round(d['value'],1), round(d['ms_per_step'],3), r['kernels']['gemm_f16_kernel']['ms'])"
UDB_GEMM_BN_AUTO=0 python bench.py --workload v1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('v1 bn256', round(d['value'],1), round(d['ms_per_step'],3), r['kernels']['gemm_f16_kernel']['ms'])"
