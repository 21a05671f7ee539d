set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "PARITY|passed|failed|Error|fused" > gpurun_out/r02_parity_gpu_fused.log; tail -45 gpurun_out/r02_parity_gpu_fused.log
for i in 1 2; do
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_fuse_$i.json 2> gpurun_out/r02_bench_fuse_$i.err; tail -c 300 gpurun_out/r02_bench_fuse_$i.json | head -c 10; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_fuse_$i.json').read().strip().splitlines()[-1]); r=d['roofline']; print('FUSED', round(d['value'],1), round(d['ms_per_step'],3), r['frac'], {k:v['ms'] for k,v in list(r['kernels'].items())[:4]})"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fuse-ln > gpurun_out/r02_bench_nofuse_$i.json 2> gpurun_out/r02_bench_nofuse_$i.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_nofuse_$i.json').read().strip().splitlines()[-1]); r=d['roofline']; print('UNFUSED', round(d['value'],1), round(d['ms_per_step'],3), r['frac'], {k:v['ms'] for k,v in list(r['kernels'].items())[:4]})"
done
tail -3 gpurun_out/r02_bench_fuse_1.err
