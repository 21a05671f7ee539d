set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "PARITY|passed|failed|Error|error|FAIL|split" > gpurun_out/r02_parity_gpu.log
tail -40 gpurun_out/r02_parity_gpu.log
python bench.py --impl torch-gpu --steps 10 --warmup 3 > gpurun_out/r02_bench_torchgpu.json 2> gpurun_out/r02_bench_torchgpu.err; cat gpurun_out/r02_bench_torchgpu.json; tail -3 gpurun_out/r02_bench_torchgpu.err
python bench.py --impl torch-gpu --workload hires --steps 5 --warmup 3 > gpurun_out/r02_bench_torchgpu_hires.json 2> gpurun_out/r02_bench_torchgpu_hires.err; cat gpurun_out/r02_bench_torchgpu_hires.json; tail -3 gpurun_out/r02_bench_torchgpu_hires.err
