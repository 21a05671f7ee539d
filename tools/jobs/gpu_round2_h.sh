set -x
mkdir -p gpurun_out
nvidia-smi topo -m | head -6
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -s 2>&1 | tail -25 > gpurun_out/r02_multigpu_test.log; cat gpurun_out/r02_multigpu_test.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_h_n1.json 2> gpurun_out/r02_bench_h_n1.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_h_n1.json').read().strip().splitlines()[-1]); print('N1', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_h_n2_p2p.json 2> gpurun_out/r02_bench_h_n2_p2p.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_h_n2_p2p.json').read().strip().splitlines()[-1]); print('N2 p2p', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['config'].get('collective','')[:80])"; tail -3 gpurun_out/r02_bench_h_n2_p2p.err
UDB_GATHER=nccl python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_h_n2_nccl.json 2> gpurun_out/r02_bench_h_n2_nccl.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_h_n2_nccl.json').read().strip().splitlines()[-1]); print('N2 nccl', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['config'].get('collective','')[:80])"; tail -3 gpurun_out/r02_bench_h_n2_nccl.err
