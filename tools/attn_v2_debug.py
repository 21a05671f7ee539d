"""Diagnostics for the attention kernel's P-in-TMEM path: with V = identity rows the output IS the probability matrix as
the PV MMA saw it, so a wrong TMEM A-operand layout shows up as a key permutation.  python tools/attn_v2_debug.py"""
import torch
from unidepth_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
d = 64
for Sq, Sk in ((128, 64), (128, 128), (128, 256)):
    q = torch.randn(Sq, d, device=dev).half()
    k = torch.randn(Sk, d, device=dev).half()
    v = torch.zeros(Sk, d, device=dev)
    idx = torch.arange(Sk, device=dev)
    v[idx, idx % 64] = 1.0 + (idx // 64).float()
    kv = torch.cat([k, v.half()], 1).contiguous()
    out = torch.empty(Sq, d, device=dev, dtype=torch.float16)
    ops.attention(q, kv, kv, out, B=1, heads=1, seq_q=Sq, seq_k=Sk, head_dim=d, k_col0=0, v_col0=d)
    torch.cuda.synchronize()
    P = torch.softmax(q.float() @ k.float().T * d ** -0.5, -1)
    ref = P @ v
    err = (out.float() - ref).abs()
    print(f"Sq{Sq} Sk{Sk}: max err {err.max().item():.3e}  finite {torch.isfinite(out).all().item()}  row-sum out {out.float().sum(1)[:4].tolist()} ref {ref.sum(1)[:4].tolist()}")
    if err.max() > 4e-3 and Sk == 64:
        o = out.float()
        # which key does each output column hold?  correlate column d of out with every column of P
        c = (o.T @ P) / (o.norm(dim=0)[:, None] * P.norm(dim=0)[None, :] + 1e-9)
        print("  column -> best matching key:", c.argmax(1).tolist())
        print("  per-row-block error:", [round(err[r:r + 32].max().item(), 4) for r in range(0, Sq, 32)])
