"""Event timeline of one attention CTA (needs a -DUDB_ATTN_TRACE variant build, see tools/build_variant.sh).
Times are clock64 deltas relative to the first softmax event."""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unidepth_b200 import ops, _cabi

lib = ctypes.CDLL(_cabi.LIB_PATH)
dev = torch.device("cuda:0")
B, H, S = 8, 16, 1611
D = H * 64
qkv = torch.randn(B * S, 3 * D, device=dev).half()
out = torch.empty(B * S, D, device=dev, dtype=torch.float16)
for _ in range(2):
    ops.attention(qkv, qkv, qkv, out, B=B, heads=H, seq_q=S, seq_k=S, head_dim=64, k_col0=D, v_col0=2 * D)
    torch.cuda.synchronize()
buf = (ctypes.c_longlong * (32 * 16))()
lib.udb_attn_trace_read(buf)
ev = [[buf[j * 16 + k] for k in range(16)] for j in range(13)]
t0 = ev[0][0]
# g0 / g1 = softmax warps 2 / 6 (first warp of each stream), mma = the issuer; c0 / c1 = 32-key chunks of the stream's half
names = {0: "g0:s_full", 1: "g0:c0", 2: "g0:c1", 3: "g0:arrived", 4: "g1:s_full", 5: "g1:c0", 6: "g1:c1", 7: "g1:arrived",
         8: "mma:p0", 9: "mma:iss0", 10: "mma:p1", 11: "mma:iss1"}
for j in range(13):
    items = sorted((ev[j][k] - t0, names[k]) for k in names if ev[j][k])
    print(f"tile {j:2d}: " + "  ".join(f"{n}@{t}" for t, n in items))
