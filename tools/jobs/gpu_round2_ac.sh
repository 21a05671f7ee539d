set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "PARITY|passed|failed|FAILED" > gpurun_out/r02_parity_gpu.log; tail -12 gpurun_out/r02_parity_gpu.log | grep -E "passed|failed|FAILED"
b() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/r02_bench_ac_$tag.json 2>gpurun_out/err_$tag.txt; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_ac_$tag.json').read().strip().splitlines()[-1]); r=d['roofline']; print('$tag', round(d['value'],1), round(d['ms_per_step'],3), r['frac'], r['kernels']['gemm_f16_kernel']['ms'], r['kernels'].get('attn_fwd_kernel',{}).get('ms'))" || tail -5 gpurun_out/err_$tag.txt; }
b pf X=1
b nopf UDB_LIB=/root/repo/variants/libudb_nopf.so
b pf_b X=1
b nopf_b UDB_LIB=/root/repo/variants/libudb_nopf.so
EXTRA="--workload hires"
b hires_pf X=1
b hires_nopf UDB_LIB=/root/repo/variants/libudb_nopf.so
EXTRA="--workload v1"
b v1_pf X=1
b v1_nopf UDB_LIB=/root/repo/variants/libudb_nopf.so
EXTRA=""
UDB_LIB=/root/repo/variants/libudb_trace.so timeout 300 python tools/attn_trace.py 2>&1 | tail -12 | head -6 | tee gpurun_out/r02_attn2_trace.txt
timeout 600 python tools/bench_kernels.py gemm 2>&1 | grep -v Warn | tee gpurun_out/r02_kernels_gemm.txt
