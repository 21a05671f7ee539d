set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_d_n1.json 2> gpurun_out/r02_bench_d_n1.err; tail -c 600 gpurun_out/r02_bench_d_n1.json; tail -3 gpurun_out/r02_bench_d_n1.err
python bench.py --workload v1 --steps 10 --warmup 3 > gpurun_out/r02_bench_v1b_n1.json 2> gpurun_out/r02_bench_v1b_n1.err; tail -c 600 gpurun_out/r02_bench_v1b_n1.json; tail -3 gpurun_out/r02_bench_v1b_n1.err
