set -x
timeout 300 python -m pytest tests/test_elementwise_gpu.py -q -s -k peer_gather 2>&1 | tail -15
timeout 600 python -m pytest tests/test_multigpu_gpu.py -q -s -k p2p 2>&1 | grep -E "rank [01]|passed|failed" | head -40
