"""Run the attention kernel a few times (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unidepth_b200 import ops
B, H, S = 8, 16, 1611
D = H * 64
qkv = torch.randn(B * S, 3 * D, device="cuda").half()
out = torch.empty(B * S, D, device="cuda", dtype=torch.float16)
for _ in range(3):
    ops.attention(qkv, qkv, qkv, out, B=B, heads=H, seq_q=S, seq_k=S, head_dim=64, k_col0=D, v_col0=2 * D)
torch.cuda.synchronize()
print("done")
