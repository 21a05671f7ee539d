set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu_gpu.py -q -s -k p2p 2>&1 | grep -E "rank|Error|error|Traceback|passed|failed|File" | head -40
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_i_$tag.json 2> gpurun_out/r02_bench_i_$tag.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_i_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))" || tail -5 gpurun_out/r02_bench_i_$tag.err; }
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_i_n1.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_i_n1.json').read().strip().splitlines()[-1]); print('N1', round(d['value'],1), round(d['ms_per_step'],3))"
run base X=1
run selfmain UDB_P2P_SELFCOPY=main
run noinput UDB_SKIP_INPUT_COPY=1
run both UDB_SKIP_INPUT_COPY=1 UDB_P2P_SELFCOPY=main
