set -x
for mode in p2p nccl; do
echo "=== FIRST=graph UDB_GATHER=$mode"
FIRST=graph UDB_GATHER=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) tools/p2p_graph_bug.py 2>&1 | grep -E "^rank|Error|error" | sort | head -40
done
