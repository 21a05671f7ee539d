"""Per-phase clock breakdown of the attention softmax warps (needs a -DUDB_ATTN_TIMING variant build):
    tools/build_variant.sh /root/repo/variants/libudb_timing.so -DUDB_ATTN_TIMING
    UDB_LIB=/root/repo/variants/libudb_timing.so python tools/attn_phases.py"""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unidepth_b200 import ops, _cabi

lib = ctypes.CDLL(_cabi.LIB_PATH)
dev = torch.device("cuda:0")
NAMES = ["wait s_full", "tmem ld S", "max+exp (+rescale)", "arrive s_free", "wait p_free", "STS P + arrive",
         "final wait"]
for (B, H, S) in [(8, 16, 1611)]:
    D = H * 64
    qkv = torch.randn(B * S, 3 * D, device=dev).half()
    out = torch.empty(B * S, D, device=dev, dtype=torch.float16)
    buf = (ctypes.c_ulonglong * 8)()
    ops.attention(qkv, qkv, qkv, out, B=B, heads=H, seq_q=S, seq_k=S, head_dim=64, k_col0=D, v_col0=2 * D)
    torch.cuda.synchronize()
    lib.udb_attn_phase_read(buf, 1)
    ops.attention(qkv, qkv, qkv, out, B=B, heads=H, seq_q=S, seq_k=S, head_dim=64, k_col0=D, v_col0=2 * D)
    torch.cuda.synchronize()
    lib.udb_attn_phase_read(buf, 1)
    tiles = buf[7]          # warp-tiles
    tot = sum(buf[k] for k in range(7))
    print(f"B{B} H{H} S{S}: {tiles} warp-tiles, {tot / tiles:.0f} clk per tile per warp")
    for k in range(7):
        print(f"  {NAMES[k]:22s} {buf[k] / tiles:8.1f} clk/tile  {100.0 * buf[k] / tot:5.1f}%")
