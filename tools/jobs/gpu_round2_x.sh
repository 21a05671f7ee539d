set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py -x -q -s 2>&1 | grep -E "attention|passed|failed|rror" | tail -24
timeout 300 python tools/attn_accuracy.py 2>&1 | grep -E "^v" 
UDB_LIB=/root/repo/variants/libudb_trace.so timeout 300 python tools/attn_trace.py 2>&1 | tail -13 | tee gpurun_out/r02_attn3_trace.txt
b() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/r02_bench_x_$tag.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_x_$tag.json').read().strip().splitlines()[-1]); r=d['roofline']; print('$tag', round(d['value'],1), round(d['ms_per_step'],3), r['frac'], r['kernels']['gemm_f16_kernel']['ms'], r['kernels']['attn_fwd_kernel']['ms'])"; }
b v3 X=1
b v2 UDB_ATTN_V=2
b v3_b X=1
b v2_b UDB_ATTN_V=2
EXTRA="--workload hires"
b hires_v3 X=1
b hires_v2 UDB_ATTN_V=2
EXTRA=""
timeout 600 python tools/bench_kernels.py attn 2>&1 | grep -v Warn | tee gpurun_out/r02_kernels_attn.txt
