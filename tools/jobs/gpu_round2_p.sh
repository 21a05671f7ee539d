set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err; tail -c 400 gpurun_out/r02_bench_final_n1.json
python bench.py --workload hires --steps 10 --warmup 3 > gpurun_out/r02_bench_final_hires.json 2> gpurun_out/r02_bench_final_hires.err; tail -c 300 gpurun_out/r02_bench_final_hires.json
python bench.py --workload v1 --steps 10 --warmup 3 > gpurun_out/r02_bench_final_v1.json 2> gpurun_out/r02_bench_final_v1.err; tail -c 300 gpurun_out/r02_bench_final_v1.json
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_step_b8.csv python tools/profile_step.py 8 default > /dev/null 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_hires_b4.csv python tools/profile_step.py 4 hires > /dev/null 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm2_f16_kernel -s 1 -c 4 -o gpurun_out/r02_gemm2 python tools/profile_step.py 8 default > /dev/null 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -c 1 -o gpurun_out/r02_attn python tools/profile_step.py 8 default > /dev/null 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -c 1 -o gpurun_out/r02_attn_hires python tools/profile_step.py 4 hires > /dev/null 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dwconv7_kernel -s 10 -c 1 -o gpurun_out/r02_dwconv python tools/profile_step.py 16 v1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
