"""The two head convolutions once each (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unidepth_b200 import ops
dev = torch.device("cuda")
for (B, H, W, C, N) in [(8, 280, 368, 128, 64), (8, 490, 644, 64, 32)]:
    x = torch.randn(B, H + 2, W + 2, C, device=dev).half()
    w = torch.randn(N, 9 * C, device=dev).half()
    bias = torch.randn(N, device=dev)
    for _ in range(2):
        if N == 64:
            ops.conv3x3_halo(x, w, bias=bias)
        else:
            ops.conv3x3_halo(x, w, bias=bias, act=ops.ACT_LEAKY, head_w=torch.randn(32, device=dev))
torch.cuda.synchronize()
print("done")
