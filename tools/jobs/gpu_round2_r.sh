set -x
mkdir -p gpurun_out
b() { tag=$1; shift; env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_r_$tag.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_r_$tag.json').read().strip().splitlines()[-1]); r=d['roofline']; print('$tag', round(d['value'],1), round(d['ms_per_step'],3), r['frac'], r['kernels']['gemm_f16_kernel']['ms'], r['kernels']['attn_fwd_kernel']['ms'])"; }
for i in 1 2; do
b base_$i X=1
b poly2_$i UDB_LIB=/root/repo/variants/libudb_poly2.so
b poly3_$i UDB_LIB=/root/repo/variants/libudb_poly3.so
b poly4_$i UDB_LIB=/root/repo/variants/libudb_poly4.so
b poly5_$i UDB_LIB=/root/repo/variants/libudb_poly5.so
done
UDB_LIB=/root/repo/variants/libudb_poly3.so python bench.py --workload hires --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('hires poly3', round(d['value'],1), round(d['ms_per_step'],3), r['kernels']['attn_fwd_kernel']['ms'])"
python bench.py --workload hires --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('hires base', round(d['value'],1), round(d['ms_per_step'],3), r['kernels']['attn_fwd_kernel']['ms'])"
UDB_LIB=/root/repo/variants/libudb_poly4.so timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
