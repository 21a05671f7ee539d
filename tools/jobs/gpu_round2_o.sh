set -x
mkdir -p gpurun_out
run() { tag=$1; n=$2; shift; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $n --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_o_$tag.json 2> gpurun_out/r02_bench_o_$tag.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_o_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))" || tail -8 gpurun_out/r02_bench_o_$tag.err; }
run nogather 2 UDB_BENCH_NOGATHER=1
run p2p_default 2 X=1
run p2p_conn32 2 CUDA_DEVICE_MAX_CONNECTIONS=32
run p2p_conn1 2 CUDA_DEVICE_MAX_CONNECTIONS=1
run nccl_conn32 2 CUDA_DEVICE_MAX_CONNECTIONS=32 UDB_GATHER=nccl
