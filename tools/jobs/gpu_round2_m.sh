set -x
mkdir -p gpurun_out
run() { tag=$1; n=$2; shift; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_m_$tag.json 2> gpurun_out/r02_bench_m_$tag.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_m_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['e2e_full']['value'],1), d['config'].get('collective','')[:60])" || tail -8 gpurun_out/r02_bench_m_$tag.err; }
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_m_n1.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_m_n1.json').read().strip().splitlines()[-1]); print('N1', round(d['value'],1), round(d['ms_per_step'],3))"
run n8_p2p 8 X=1
run n8_nccl 8 UDB_GATHER=nccl
run n4_p2p 4 X=1
