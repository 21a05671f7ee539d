set -x
mkdir -p gpurun_out
timeout 120 ./variants/mma_probe 2>&1 | head -24 | tee gpurun_out/r02_mma_probe.txt
timeout 600 python -m pytest tests/test_attention_gpu.py -x -q 2>&1 | tail -3
UDB_LIB=/root/repo/variants/libudb_trace.so timeout 300 python tools/attn_trace.py 2>&1 | tail -13 | tee gpurun_out/r02_attn3_trace.txt
b() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/r02_bench_z_$tag.json 2>gpurun_out/err_$tag.txt; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_z_$tag.json').read().strip().splitlines()[-1]); r=d['roofline']; print('$tag', round(d['value'],1), round(d['ms_per_step'],3), r['frac'], r['kernels']['gemm_f16_kernel']['ms'], r['kernels']['attn_fwd_kernel']['ms'])" || tail -5 gpurun_out/err_$tag.txt; }
b v3 X=1
b v2 UDB_ATTN_V=2
EXTRA="--workload hires"
b hires_v3 X=1
b hires_v2 UDB_ATTN_V=2
b hires_v3_b X=1
EXTRA=""
