set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "PARITY|passed|failed|FAILED" > gpurun_out/r02_parity_gpu_attn2.log; tail -12 gpurun_out/r02_parity_gpu_attn2.log
b() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/r02_bench_w_$tag.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_w_$tag.json').read().strip().splitlines()[-1]); r=d['roofline']; print('$tag', round(d['value'],1), round(d['ms_per_step'],3), r['frac'], r['kernels']['gemm_f16_kernel']['ms'], r['kernels']['attn_fwd_kernel']['ms'])"; }
b base X=1
b poly3 UDB_LIB=/root/repo/variants/libudb_poly3.so
b poly6 UDB_LIB=/root/repo/variants/libudb_poly6.so
b ns2 UDB_LIB=/root/repo/variants/libudb_ns2.so
b base_b X=1
b old_kernel UDB_ATTN_V=1
UDB_LIB=/root/repo/variants/libudb_trace.so timeout 300 python tools/attn_trace.py 2>&1 | tail -14 | tee gpurun_out/r02_attn2_trace.txt
timeout 600 python tools/bench_kernels.py attn 2>&1 | grep -v Warn | tee gpurun_out/r02_kernels_attn.txt
timeout 600 python tools/bench_kernels.py gemm 2>&1 | grep -v Warn | tee gpurun_out/r02_kernels_gemm.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -c 1 -o gpurun_out/r02_attn2 python tools/profile_attn.py > /dev/null 2>&1
