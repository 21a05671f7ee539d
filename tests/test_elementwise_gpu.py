"""HBM-bound kernels vs the torch ops the reference calls (fp32)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def _err(got, ref, name):
    e = (got.float() - ref.float()).abs().max().item()
    print(f"{name}: max abs err {e:.3e} (ref max {ref.abs().max().item():.3e})")
    return e


@pytest.mark.parametrize("dim", [128, 384, 512, 1024])
def test_layernorm(dim):
    from unidepth_b200 import ops
    dev = _dev()
    torch.manual_seed(dim)
    x = torch.randn(1003, dim, device=dev) * 3 + 0.5
    w, b = torch.randn(dim, device=dev), torch.randn(dim, device=dev)
    ref = F.layer_norm(x, (dim,), w, b, 1e-6)
    assert _err(ops.layernorm(x, w, b, 1e-6, out_dtype=torch.float32), ref, "ln f32") < 2e-5
    assert _err(ops.layernorm(x, w, b, 1e-6, out_dtype=torch.float16), ref, "ln f16") < 8e-3
    assert _err(ops.layernorm(x.half(), w, b, 1e-5, out_dtype=torch.float32),
                F.layer_norm(x.half().float(), (dim,), w, b, 1e-5), "ln f16 in") < 2e-5
    if dim in (128,):
        for d2 in (64, 128, 256):
            xs = (torch.randn(100003, d2, device=dev) * 2 + 0.3).half()
            w2, b2 = torch.randn(d2, device=dev), torch.randn(d2, device=dev)
            assert _err(ops.layernorm(xs, w2, b2, 1e-5, out_dtype=torch.float16),
                        F.layer_norm(xs.float(), (d2,), w2, b2, 1e-5), f"ln f16->f16 small {d2}") < 8e-3
    # drop the cls row of each image: rows (b, 1 + n)
    Bn, T = 7, 143
    xt = torch.randn(Bn * T, dim, device=dev)
    got = ops.layernorm(xt, w, b, 1e-5, out_dtype=torch.float32, rows=Bn * (T - 1), rows_per_group=T - 1,
                        group_stride=T, row_offset=1)
    ref = F.layer_norm(xt.view(Bn, T, dim)[:, 1:], (dim,), w, b, 1e-5).reshape(-1, dim)
    assert _err(got, ref, "ln row map") < 2e-5


@pytest.mark.parametrize("shape,pads,net", [((2, 480, 640), (0, 0, 0, 0), (490, 644)), ((1, 96, 288), (0, 0, 9, 10), (406, 1022)),
                                            ((1, 200, 90), (5, 5, 0, 0), (644, 322))])
def test_preprocess_patchify(shape, pads, net):
    from unidepth_b200 import ops
    dev = _dev()
    B, H, W = shape
    g = torch.Generator().manual_seed(5)
    rgb = torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, generator=g).to(dev)
    mean = torch.tensor([0.485, 0.456, 0.406], device=dev).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device=dev).view(1, 3, 1, 1)
    x = (rgb.float() / 255.0 - mean) / std
    x = F.pad(x, pads, value=0.0)
    x = F.interpolate(x, size=net, mode="bilinear", align_corners=False)
    gh, gw = net[0] // 14, net[1] // 14
    ref = F.unfold(x, kernel_size=14, stride=14).transpose(1, 2).reshape(B * gh * gw, 588)
    patches = torch.empty(B * gh * gw, 640, device=dev, dtype=torch.float16)
    ops.preprocess_patchify(rgb, pads, net, patches)
    assert _err(patches[:, :588], ref, "patchify") < 2e-3
    assert patches[:, 588:].abs().max().item() == 0


def test_posembed_bicubic():
    from unidepth_b200 import ops
    dev = _dev()
    torch.manual_seed(0)
    m, dim = 37, 256
    grid = torch.randn(m * m, dim, device=dev)
    for gh, gw in [(35, 46), (46, 68), (29, 73)]:
        ref = F.interpolate(grid.view(1, m, m, dim).permute(0, 3, 1, 2), size=(gh, gw), mode="bicubic",
                            antialias=False).permute(0, 2, 3, 1).reshape(gh * gw, dim)
        assert _err(ops.posembed_bicubic(grid, m, dim, gh, gw), ref, f"bicubic {gh}x{gw}") < 5e-6


def test_small_linear_and_attn4():
    from unidepth_b200 import ops
    dev = _dev()
    torch.manual_seed(0)
    M, K, N = 32, 512, 2048
    x, w, b = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev) / K ** 0.5, torch.randn(N, device=dev)
    gamma, resid = torch.rand(N, device=dev), torch.randn(M, N, device=dev)
    ref = resid + gamma * F.gelu(x @ w.t() + b)
    assert _err(ops.small_linear(x, w, b, act=ops.ACT_GELU, gamma=gamma, resid=resid), ref, "small linear") < 1e-4
    w1 = torch.randn(1, K, device=dev) / K ** 0.5
    assert _err(ops.small_linear(x, w1, None), x @ w1.t(), "small linear N=1") < 1e-5
    Bn, Cc, heads = 5, 512, 8
    q, kv, pos = torch.randn(Bn, 4, Cc, device=dev), torch.randn(Bn, 4, 2 * Cc, device=dev), torch.randn(4, Cc, device=dev)
    hd = lambda t: t.view(Bn, 4, heads, Cc // heads).transpose(1, 2)
    ref = F.scaled_dot_product_attention(hd(q + pos), hd(kv[..., :Cc]), hd(kv[..., Cc:])).transpose(1, 2).reshape(Bn, 4, Cc)
    assert _err(ops.camera_attn4(q, kv, pos, Bn, Cc, heads), ref, "attn4") < 1e-5


def _rays(intr, hh, ww):
    import unidepth_oracle as O
    return O.rays_from_intrinsics(intr, hh, ww)


def test_camera_rays_embed_postprocess():
    from unidepth_b200 import ops
    import unidepth_oracle as O
    dev = _dev()
    torch.manual_seed(0)
    B, nh, nw = 2, 490, 644
    x = torch.randn(B, 4) * 0.3
    diag = (nh ** 2 + nw ** 2) ** 0.5
    intr_ref = torch.stack([x[:, 0].exp() * 0.7 * diag, x[:, 1].exp() * 0.7 * diag, x[:, 2].sigmoid() * nw, x[:, 3].sigmoid() * nh], 1)
    intr4, k_net, k_out = ops.camera_intrinsics(x.to(dev), B, (nh, nw), 1.02, 3, 0)
    assert _err(intr4.cpu(), intr_ref, "intr4") < 1e-3 * 1e-1
    kmat, rays = O.rays_from_intrinsics(intr4.cpu(), nh, nw)
    assert _err(k_net.cpu(), kmat, "K") == 0
    gh, gw = nh // 14, nw // 14
    ref = O.embed_rays(rays, (nh, nw), (gh, gw), 512)
    scales = (2.0 ** torch.linspace(0.0, math.log2(max(gh, gw) // 2), steps=256)).to(dev)
    emb = ops.ray_embed(intr4, scales, B, (nh, nw), (gh, gw), out_dtype=torch.float32)
    assert _err(emb.cpu(), ref.reshape(-1, 512), "ray embedding") < 2e-3
    emb2 = ops.ray_embed(intr4, scales, B, (nh, nw), (gh, gw), out_dtype=torch.float32, rays_in=rays.to(dev).contiguous())
    assert _err(emb2.cpu(), ref.reshape(-1, 512), "ray embedding (rays_in)") < 2e-3
    # postprocess
    radius = torch.rand(B, nh, nw) * 5 + 1
    conf = torch.rand(B, nh, nw) + 0.5
    H, W, ph, pw, pl, pt = 480, 634, 480, 640, 3, 0
    outs = ops.postprocess(radius.to(dev), conf.to(dev), intr4, B, (nh, nw), (ph, pw), pl, pt, (H, W))
    rays_map = rays.transpose(1, 2).reshape(B, 3, nh, nw)
    pts = O._post(rays_map * radius.unsqueeze(1), (ph, pw), (pl, pw - W - pl, pt, ph - H - pt))
    rys = O._post(rays_map, (ph, pw), (pl, pw - W - pl, pt, ph - H - pt))
    cf = O._post(conf.unsqueeze(1), (ph, pw), (pl, pw - W - pl, pt, ph - H - pt))
    assert _err(outs["points"].cpu(), pts, "points") < 2e-5
    assert _err(outs["confidence"].cpu(), cf, "confidence") < 2e-6
    assert _err(outs["depth"].cpu(), pts[:, -1:], "depth") < 2e-5
    assert _err(outs["radius"].cpu(), pts.norm(dim=1, keepdim=True), "radius") < 2e-5
    assert _err(outs["rays"].cpu(), rys / rys.norm(dim=1, keepdim=True).clip(min=1e-5), "rays") < 2e-6


def test_resamplers():
    from unidepth_b200 import ops
    dev = _dev()
    torch.manual_seed(0)
    x = torch.randn(2, 64, 35, 46, device=dev).half()
    xn = x.permute(0, 2, 3, 1).contiguous()
    ref = F.interpolate(x.float(), scale_factor=2, mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    assert _err(ops.upsample2x(xn), ref, "upsample2x") < 4e-3
    ref = F.interpolate(x.float(), size=(61, 80), mode="bilinear", align_corners=True)
    got = ops.resize_ac_pad(xn, 61, 80, 0)
    assert _err(got, ref.permute(0, 2, 3, 1), "resize ac") < 4e-3
    got = ops.resize_ac_pad(xn, 61, 80, 1)
    refp = F.pad(ref, (1, 1, 1, 1), mode="reflect").permute(0, 2, 3, 1)
    assert _err(got, refp, "resize ac + reflect pad") < 4e-3


def test_layernorm_zero_padded_channels():
    """ViT-B's 96-channel map is stored as 128 channels (32 zeros): statistics over the first 96 only,
    padded outputs stay zero (weight 0 there)."""
    from unidepth_b200 import ops
    dev = _dev()
    torch.manual_seed(5)
    for dim, valid in ((128, 96), (256, 192), (64, 48)):
        x = torch.zeros(50003, dim, device=dev)
        x[:, :valid] = torch.randn(50003, valid, device=dev) * 2 + 0.7
        x = x.half()
        w = torch.zeros(dim, device=dev)
        w[:valid] = 1.0
        b = torch.zeros(dim, device=dev)
        got = ops.layernorm(x, w, b, 1e-5, out_dtype=torch.float16, dim_valid=valid)
        ref = F.layer_norm(x[:, :valid].float(), (valid,), None, None, 1e-5)
        assert _err(got[:, :valid], ref, f"ln {valid} of {dim}") < 8e-3
        assert float(got[:, valid:].abs().max()) == 0.0


def test_peer_gather_single_rank():
    """parallel.PeerGather with a one-rank NCCL group: the send-slot views alias the CUDA-IPC buffer (writes through them are
    what start() sends), slots rotate, outputs written elsewhere are staged by one copy, permuted outputs come back with the
    same values.  (The cross-process part -- IPC mapping and the device barrier -- is tests/test_multigpu_gpu.py.)"""
    import os
    import torch.distributed as dist
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from unidepth_b200 import parallel
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29655")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    try:
        g = torch.Generator().manual_seed(0)
        mk = lambda: {"intrinsics": torch.randn(3, 3, 3, generator=g).to(dev), "depth": torch.randn(3, 1, 20, 30, generator=g).to(dev),
                      "depth_features": torch.randn(3, 5, 7, 16, generator=g).to(dev).permute(0, 3, 1, 2)}
        a = mk()
        pg = parallel.PeerGather(a, dev)
        for step in range(5):
            src = mk()
            if step % 2:            # produced in place: write through the views
                v = pg.views()
                for k in src:
                    v[k].copy_(src[k])
                got = pg.start(v).wait()
            else:                   # produced elsewhere: start() stages it
                got = pg.start(src).wait()
            torch.cuda.synchronize()
            for k in src:
                assert got[k].shape == src[k].shape and got[k].stride() == src[k].stride() and torch.equal(got[k], src[k]), (step, k)
        assert not pg.timed_out()
    finally:
        if created:
            dist.destroy_process_group()
