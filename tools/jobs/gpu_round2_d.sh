set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_v1_gpu.py -q 2>&1 | tail -3
python bench.py --workload v1 --steps 10 --warmup 3 > gpurun_out/r02_bench_v1_n1.json 2> gpurun_out/r02_bench_v1_n1.err; tail -c 3500 gpurun_out/r02_bench_v1_n1.json; tail -3 gpurun_out/r02_bench_v1_n1.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_b_n1.json 2> gpurun_out/r02_bench_b_n1.err; tail -c 2500 gpurun_out/r02_bench_b_n1.json; tail -3 gpurun_out/r02_bench_b_n1.err
python bench.py --impl torch-gpu --workload v1 --steps 5 --warmup 3 > gpurun_out/r02_bench_torchgpu_v1.json 2> gpurun_out/r02_bench_torchgpu_v1.err; cat gpurun_out/r02_bench_torchgpu_v1.json; tail -3 gpurun_out/r02_bench_torchgpu_v1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_ncu_launches_v1_b16.csv python bench.py --workload v1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_v1.log 2>&1; tail -2 gpurun_out/ncu_v1.log | cut -c1-300
