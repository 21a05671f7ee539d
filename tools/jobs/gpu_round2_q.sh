set -x
mkdir -p gpurun_out
for v in pack packdirect; do UDB_LIB=/root/repo/variants/libudb_$v.so timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_infer_parity_gpu.py -m gpu -q -x -k "gemm or golden_vitl or shallow_vitl or against_reference" 2>&1 | tail -2; done
for v in poly4 poly8; do UDB_LIB=/root/repo/variants/libudb_$v.so timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_infer_parity_gpu.py -m gpu -q -x -k "attention or golden_vitl or shallow_vitl or against_reference" 2>&1 | tail -2; done
b() { tag=$1; shift; env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_q_$tag.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_q_$tag.json').read().strip().splitlines()[-1]); r=d['roofline']; print('$tag', round(d['value'],1), round(d['ms_per_step'],3), r['frac'], r['kernels']['gemm_f16_kernel']['ms'], r['kernels']['attn_fwd_kernel']['ms'])"; }
for i in 1 2; do
b base_$i X=1
b pack_$i UDB_LIB=/root/repo/variants/libudb_pack.so
b packdirect_$i UDB_LIB=/root/repo/variants/libudb_packdirect.so
b poly4_$i UDB_LIB=/root/repo/variants/libudb_poly4.so
b poly8_$i UDB_LIB=/root/repo/variants/libudb_poly8.so
done
