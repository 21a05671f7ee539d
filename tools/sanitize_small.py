"""Small end-to-end and per-kernel invocations for compute-sanitizer (memcheck / racecheck are 10-100x slower, so small shapes):
  compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_small.py
Covers: the tcgen05 GEMM (pair + single-CTA flavours, epilogue variants), attention (ragged key counts), LayerNorm, the camera-head
linears and one shallow ViT-S `infer` through the C engine (every kernel of the V2 path at least once)."""
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from unidepth_b200 import UniDepthV2, ops  # noqa: E402
from unidepth_b200.synthetic import synthetic_state_dict  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
# attention: full tiles, ragged last tile, second stream without keys
for (B, H, Sq, Sk) in ((1, 2, 200, 200), (1, 1, 100, 40), (2, 1, 130, 300)):
    D = H * 64
    q = torch.randn(B * Sq, D, device=dev).half()
    kv = torch.randn(B * Sk, 2 * D, device=dev).half()
    out = torch.empty(B * Sq, D, device=dev, dtype=torch.float16)
    ops.attention(q, kv, kv, out, B=B, heads=H, seq_q=Sq, seq_k=Sk, head_dim=64, k_col0=0, v_col0=D)
# GEMM: pair kernel with bias / GELU, f32 residual in place, narrow single-CTA tile
a = torch.randn(300, 256, device=dev).half()
w = torch.randn(512, 256, device=dev).half()
ops.gemm(a, w, bias=torch.randn(512, device=dev), act=ops.ACT_GELU, out=torch.empty(300, 512, device=dev, dtype=torch.float16))
x = torch.randn(300, 512, device=dev)
ops.gemm(a, w, bias=torch.randn(512, device=dev), gamma=torch.rand(512, device=dev), resid=x, out=x)
w64 = torch.randn(64, 256, device=dev).half()
ops.gemm(a, w64, out=torch.empty(300, 64, device=dev, dtype=torch.float16))
ops.layernorm(x, torch.randn(512, device=dev), torch.randn(512, device=dev), 1e-6, out=torch.empty(300, 512, device=dev, dtype=torch.float16))
torch.cuda.synchronize()
# one shallow ViT-S infer (eager and graph replay)
cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "config_v2_vits14.json")))
cfg["model"]["pixel_encoder"]["arch_override"] = {"depth": 4}
cfg["model"]["pixel_encoder"]["output_idx"] = [1, 2, 3, 4]
m = UniDepthV2(copy.deepcopy(cfg))
m.load_state_dict(synthetic_state_dict(cfg, 0, device="cuda"), strict=True)
m = m.to(dev).eval()
rgb = torch.randint(0, 256, (2, 3, 112, 160), dtype=torch.uint8, device=dev)
m.use_cuda_graph = False
o1 = m.infer(rgb)
m.use_cuda_graph = True
o2 = m.infer(rgb)
o3 = m.infer(rgb)
torch.cuda.synchronize()
assert torch.equal(o2["depth"], o3["depth"]) and torch.isfinite(o1["depth"]).all()
print("sanitize_small: done")
