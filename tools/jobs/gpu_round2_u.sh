set -x
mkdir -p gpurun_out
timeout 120 ./variants/tmem_probe 2>&1 | tee gpurun_out/r02_tmem_probe.txt
UDB_LIB=/root/repo/variants/libudb_trace.so timeout 300 python tools/attn_trace.py 2>&1 | tail -14 | tee gpurun_out/r02_attn2_trace.txt
b() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/r02_bench_u_$tag.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_u_$tag.json').read().strip().splitlines()[-1]); r=d['roofline']; print('$tag', round(d['value'],1), round(d['ms_per_step'],3), r['frac'], r['kernels']['gemm_f16_kernel']['ms'], r['kernels']['attn_fwd_kernel']['ms'])"; }
b base X=1
b ns2 UDB_LIB=/root/repo/variants/libudb_ns2.so
b p3 UDB_LIB=/root/repo/variants/libudb_p3.so
b p6 UDB_LIB=/root/repo/variants/libudb_p6.so
b p0 UDB_LIB=/root/repo/variants/libudb_p0.so
b base2 X=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -c 1 -o gpurun_out/r02_attn2 python tools/profile_attn.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
