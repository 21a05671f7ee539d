"""Mean / max error of the attention kernel against a float64 reference on the same f16 inputs (run once per kernel version:
UDB_LIB=<variant library>).  Inputs: N(0,1) q/k/v and a 'peaky' set (q scaled by 3) closer to trained attention maps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unidepth_b200 import ops

dev = torch.device("cuda:0")
B, H, S, d = 2, 16, 1611, 64
D = H * d
for name, gain in (("normal", 1.0), ("peaky", 3.0), ("very peaky", 6.0)):
    torch.manual_seed(1)
    qkv = torch.randn(B * S, 3 * D, device=dev)
    qkv[:, :D] *= gain
    qkv = qkv.half()
    out = torch.empty(B * S, D, device=dev, dtype=torch.float16)
    ops.attention(qkv, qkv, qkv, out, B=B, heads=H, seq_q=S, seq_k=S, head_dim=d, q_col0=0, k_col0=D, v_col0=2 * D)
    q, k, v = qkv.view(B, S, 3, H, d).permute(2, 0, 3, 1, 4).double()
    ref = (torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5, -1) @ v).permute(0, 2, 1, 3).reshape(B * S, D)
    e = (out.double() - ref).abs()
    r16 = (ref.half().double() - ref).abs()          # the output rounding alone
    print(f"{name:10s}: mean abs err {e.mean().item():.4e}  max {e.max().item():.3e}  rms {e.pow(2).mean().sqrt().item():.4e}"
          f"   (f16 rounding of the exact result alone: mean {r16.mean().item():.4e})  ref rms {ref.pow(2).mean().sqrt().item():.3e}")
