set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_elementwise_gpu.py -q -k peer_gather 2>&1 | tail -2
timeout 600 python -m pytest tests/test_multigpu_gpu.py -q 2>&1 | tail -2
run() { tag=$1; n=$2; shift; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $n --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n_$tag.json 2> gpurun_out/r02_bench_n_$tag.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))" || tail -8 gpurun_out/r02_bench_n_$tag.err; }
python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n_n1.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n_n1.json').read().strip().splitlines()[-1]); print('N1 gpu0', round(d['value'],1), round(d['ms_per_step'],3))"
CUDA_VISIBLE_DEVICES=1 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n_n1b.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n_n1b.json').read().strip().splitlines()[-1]); print('N1 gpu1', round(d['value'],1), round(d['ms_per_step'],3))"
run nogather 2 UDB_BENCH_NOGATHER=1
run p2p 2 X=1
