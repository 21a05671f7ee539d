set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "PARITY|passed|failed|Error|error|FAIL|smoke" > gpurun_out/r02_parity_gpu.log
tail -5 gpurun_out/r02_parity_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_a_n1.json 2> gpurun_out/r02_bench_a_n1.err; tail -c 3000 gpurun_out/r02_bench_a_n1.json
python bench.py --workload hires --steps 10 --warmup 3 > gpurun_out/r02_bench_a_hires.json 2> gpurun_out/r02_bench_a_hires.err; tail -c 2500 gpurun_out/r02_bench_a_hires.json
python bench.py --impl torch-gpu --steps 10 --warmup 3 > gpurun_out/r02_bench_torchgpu.json 2> gpurun_out/r02_bench_torchgpu.err; cat gpurun_out/r02_bench_torchgpu.json; tail -3 gpurun_out/r02_bench_torchgpu.err
python bench.py --impl torch-gpu --workload hires --steps 5 --warmup 3 > gpurun_out/r02_bench_torchgpu_hires.json 2> gpurun_out/r02_bench_torchgpu_hires.err; cat gpurun_out/r02_bench_torchgpu_hires.json
