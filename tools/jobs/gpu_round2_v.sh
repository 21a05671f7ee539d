set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py -x -q 2>&1 | tail -3
(UDB_ATTN_V=1 timeout 300 python tools/attn_accuracy.py; UDB_ATTN_V=2 timeout 300 python tools/attn_accuracy.py) 2>&1 | grep -E "^v[12]" | tee gpurun_out/r02_attn_accuracy.txt
UDB_LIB=/root/repo/variants/libudb_trace.so timeout 300 python tools/attn_trace.py 2>&1 | tail -14 | tee gpurun_out/r02_attn2_trace_b.txt
b() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/r02_bench_v_$tag.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_v_$tag.json').read().strip().splitlines()[-1]); r=d['roofline']; print('$tag', round(d['value'],1), round(d['ms_per_step'],3), r['frac'], r['kernels']['gemm_f16_kernel']['ms'], r['kernels']['attn_fwd_kernel']['ms'])"; }
b p1i2 X=1
b p0i1 UDB_LIB=/root/repo/variants/libudb_p0i1.so
b p1i1 UDB_LIB=/root/repo/variants/libudb_p1i1.so
b p0i2 UDB_LIB=/root/repo/variants/libudb_p0i2.so
b p1i2_b X=1
EXTRA="--workload hires"
b hires_p1i2 X=1
b hires_p0i1 UDB_LIB=/root/repo/variants/libudb_p0i1.so
