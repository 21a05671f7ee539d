set -x
mkdir -p gpurun_out
timeout 300 python tools/attn_v2_debug.py 2>&1 | tail -12
timeout 600 python -m pytest tests/test_attention_gpu.py -x -q -s 2>&1 | grep -E "attention|passed|failed|Error|error" | tail -30
b() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/r02_bench_t_$tag.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_t_$tag.json').read().strip().splitlines()[-1]); r=d['roofline']; print('$tag', round(d['value'],1), round(d['ms_per_step'],3), r['frac'], r['kernels']['gemm_f16_kernel']['ms'], r['kernels']['attn_fwd_kernel']['ms'])"; }
for i in 1 2; do
b v1_$i UDB_ATTN_V=1
b v2_$i UDB_ATTN_V=2
done
EXTRA="--workload hires"
b hires_v1 UDB_ATTN_V=1
b hires_v2 UDB_ATTN_V=2
EXTRA=""
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "PARITY|passed|failed|FAILED" > gpurun_out/r02_parity_gpu_attn2.log; tail -8 gpurun_out/r02_parity_gpu_attn2.log
