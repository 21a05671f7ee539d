set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_v1_gpu.py -q -s -x -k "not infer and not batch" 2>&1 | tail -40 > gpurun_out/r02_v1_units.log; cat gpurun_out/r02_v1_units.log
timeout 600 python tools/v1_debug_taps.py 2>&1 | tail -30 > gpurun_out/r02_v1_taps.log; cat gpurun_out/r02_v1_taps.log
timeout 900 python -m pytest tests/test_v1_gpu.py -q -s -k "infer or batch" 2>&1 | grep -E "V1PARITY|passed|failed|Error|assert" | head -30 > gpurun_out/r02_v1_e2e.log; cat gpurun_out/r02_v1_e2e.log
