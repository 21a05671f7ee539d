"""Fused attention kernel vs torch softmax(QK^T/sqrt(d))V in fp32 on the same f16 inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(q, k, v):
    q, k, v = q.float(), k.float(), v.float()
    a = torch.softmax(q @ k.transpose(-1, -2) * q.shape[-1] ** -0.5, dim=-1)
    return a @ v


@pytest.mark.parametrize("B,H,Sq,Sk", [(1, 1, 128, 128), (1, 2, 128, 256), (2, 3, 200, 200), (2, 16, 1611, 1611),
                                       (1, 8, 1610, 1610), (1, 4, 77, 300),
                                       # key counts around the 32-key chunk / 64-key half / 128-key tile edges: the second
                                       # softmax stream of the kernel has no valid key at all when Sk <= 64
                                       (1, 2, 100, 40), (1, 1, 128, 64), (2, 2, 130, 65), (1, 2, 64, 97), (1, 1, 50, 129),
                                       (1, 1, 256, 192), (1, 2, 300, 3129)])
def test_attention_fused_qkv_layout(B, H, Sq, Sk):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from unidepth_b200 import ops
    dev = torch.device("cuda:0")
    torch.manual_seed(Sq + H)
    d = 64
    D = H * d
    if Sq == Sk:
        qkv = torch.randn(B * Sq, 3 * D, device=dev).half()      # [3][H][d] column layout (ViT qkv)
        out = torch.empty(B * Sq, D, device=dev, dtype=torch.float16)
        ops.attention(qkv, qkv, qkv, out, B=B, heads=H, seq_q=Sq, seq_k=Sk, head_dim=d, q_col0=0, k_col0=D, v_col0=2 * D)
        q, k, v = qkv.view(B, Sq, 3, H, d).permute(2, 0, 3, 1, 4)
    else:
        qm = torch.randn(B * Sq, D, device=dev).half()
        kv = torch.randn(B * Sk, 2 * D, device=dev).half()       # [k(H d); v(H d)] (decoder kv)
        out = torch.empty(B * Sq, D, device=dev, dtype=torch.float16)
        ops.attention(qm, kv, kv, out, B=B, heads=H, seq_q=Sq, seq_k=Sk, head_dim=d, k_col0=0, v_col0=D)
        q = qm.view(B, Sq, H, d).permute(0, 2, 1, 3)
        k, v = kv.view(B, Sk, 2, H, d).permute(2, 0, 3, 1, 4)
    ref = _ref(q, k, v).permute(0, 2, 1, 3).reshape(B * Sq, D)
    err = (out.float() - ref).abs().max().item()
    print(f"attention B{B} H{H} Sq{Sq} Sk{Sk}: max abs err {err:.3e} (ref max {ref.abs().max().item():.3e})")
    assert err < 4e-3


@pytest.mark.parametrize("Sq,Sk,gain", [(300, 700, 3.0), (128, 1000, 5.0), (257, 130, 4.0)])
def test_attention_peaky_growing_logits_exercise_rescale(Sq, Sk, gain):
    """Logits with a large spread whose maxima keep growing along the key axis: the speculative
    softmax is rejected in most tiles and halves, so the redo path, the lazy rescale of O and the
    rescale of the already packed first half are all exercised (they are rare with N(0,1) inputs)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from unidepth_b200 import ops
    dev = torch.device("cuda:0")
    torch.manual_seed(Sq * 7 + Sk)
    B, H, d = 2, 2, 64
    D = H * d
    qm = (torch.randn(B * Sq, D, device=dev) * 2.0).half()
    ramp = (1.0 + gain * torch.arange(Sk, device=dev).float() / Sk).repeat(B).unsqueeze(1)
    kv = torch.randn(B * Sk, 2 * D, device=dev)
    kv[:, :D] *= ramp
    kv = kv.half()
    out = torch.empty(B * Sq, D, device=dev, dtype=torch.float16)
    ops.attention(qm, kv, kv, out, B=B, heads=H, seq_q=Sq, seq_k=Sk, head_dim=d, k_col0=0, v_col0=D)
    q = qm.view(B, Sq, H, d).permute(0, 2, 1, 3)
    k, v = kv.view(B, Sk, 2, H, d).permute(2, 0, 3, 1, 4)
    ref = _ref(q, k, v).permute(0, 2, 1, 3).reshape(B * Sq, D)
    assert torch.isfinite(out).all()
    err = (out.float() - ref).abs().max().item()
    print(f"peaky attention Sq{Sq} Sk{Sk} gain {gain}: max abs err {err:.3e} (ref max {ref.abs().max().item():.3e})")
    assert err < 8e-3
