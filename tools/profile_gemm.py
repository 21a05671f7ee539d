"""The three encoder GEMM flavours once each (for ncu): qkv (plain f16 out), fc1 (GELU), proj
(gamma * . + f32 residual in place)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unidepth_b200 import ops
dev = torch.device("cuda")
M = 12888
a = torch.randn(M, 1024, device=dev).half()
a4 = torch.randn(M, 4096, device=dev).half()
w_qkv = torch.randn(3072, 1024, device=dev).half()
w_fc1 = torch.randn(4096, 1024, device=dev).half()
w_fc2 = torch.randn(1024, 4096, device=dev).half()
w_proj = torch.randn(1024, 1024, device=dev).half()
b3, b4, b1 = torch.randn(3072, device=dev), torch.randn(4096, device=dev), torch.randn(1024, device=dev)
gamma = torch.rand(1024, device=dev)
x = torch.randn(M, 1024, device=dev)
o3 = torch.empty(M, 3072, device=dev, dtype=torch.float16)
o4 = torch.empty(M, 4096, device=dev, dtype=torch.float16)
for _ in range(2):
    ops.gemm(a, w_qkv, bias=b3, out=o3)
    ops.gemm(a, w_fc1, bias=b4, act=ops.ACT_GELU, out=o4)
    ops.gemm(a, w_proj, bias=b1, gamma=gamma, resid=x, out=x)
    ops.gemm(a4, w_fc2, bias=b1, gamma=gamma, resid=x, out=x)
torch.cuda.synchronize()
print("done")
