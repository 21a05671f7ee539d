"""CPU oracle for the UniDepthV2.infer() hot path  --  TEST INFRASTRUCTURE ONLY.

A functional fp32 restatement (plain torch ops on a flat state-dict) of the reference's
inference forward.  It exists so that parity can be checked on a machine that does not have
/root/reference (the GPU box).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product package
(unidepth_b200/) never does and has no CPU fallback.

Pinning: tests/golden/*.npz hold outputs of the *unmodified* reference (imported from
/root/reference with oracle/ref_shims, see oracle/make_golden.py) on seeded weights; the
`not gpu` test-suite checks this restatement against them (tests/test_oracle_golden.py).

Every function cites the reference file:line (relative to /root/reference) it restates.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # unidepth/utils/constants.py:12
IMAGENET_STD = (0.229, 0.224, 0.225)   # unidepth/utils/constants.py:13

# arch table: unidepth/models/backbones/dinov2.py:388-427 (vit_small/base/large) and the
# factories' default taps unidepth/models/encoder.py:139-193
VIT_ARCH = {
    "dinov2_vits14": dict(embed_dim=384, depth=12, num_heads=6),
    "dinov2_vitb14": dict(embed_dim=768, depth=12, num_heads=12),
    "dinov2_vitl14": dict(embed_dim=1024, depth=24, num_heads=16),
}
VIT_DEFAULT_TAPS = {
    "dinov2_vits14": [3, 6, 9, 12],
    "dinov2_vitb14": [3, 6, 9, 12],
    "dinov2_vitl14": [5, 12, 18, 24],
}
PATCH = 14


class ModelSpec:
    """Shape hyper-parameters pulled out of a reference config dict
    (configs/config_v2_vit*.json; unidepthv2.py:418-460, unidepthv2/decoder.py:470-524)."""

    def __init__(self, config: dict):
        enc = config["model"]["pixel_encoder"]
        dec = config["model"]["pixel_decoder"]
        name = enc["name"]
        arch = dict(VIT_ARCH[name])
        # test-only override so that small/odd encoders can be described in a config
        arch.update(enc.get("arch_override", {}))
        self.name = name
        self.embed_dim = arch["embed_dim"]
        self.depth = arch["depth"]
        self.enc_heads = arch["num_heads"]
        self.taps = list(enc.get("output_idx", VIT_DEFAULT_TAPS[name]))  # 1-based block idx
        self.hidden = dec["hidden_dim"]
        self.dec_heads = config["model"]["num_heads"]
        self.expansion = config["model"]["expansion"]
        self.dec_depths = list(dec["depths"])
        self.out_dim = dec["out_dim"]
        self.kernel_size = dec.get("kernel_size", 7)
        self.use_norm = bool(enc.get("use_norm", False))
        sc = config["data"]["augmentations"]["shape_constraints"]
        self.ratio_bounds = tuple(sc["ratio_bounds"])
        self.pixels_bounds = (sc["pixels_min"], sc["pixels_max"])


# --------------------------------------------------------------------------------------
# a1: shape arithmetic (pure Python floats/ints)
# --------------------------------------------------------------------------------------
def get_paddings(original_shape, aspect_ratio_range):
    """unidepthv2.py:36-58."""
    h, w = original_shape
    ratio = w / h
    lo, hi = aspect_ratio_range
    target = min(hi, max(lo, ratio))
    if ratio > target:  # too wide -> pad top/bottom
        h_new, w_new = int(w / target), w
        pt = (h_new - h) // 2
        return (0, 0, pt, h_new - h - pt), (h_new, w_new)
    h_new, w_new = h, int(h * target)
    pl = (w_new - w) // 2
    return (pl, w_new - w - pl, 0, 0), (h_new, w_new)


def get_resize_factor(original_shape, pixels_range, shape_multiplier=14):
    """unidepthv2.py:61-77."""
    h, w = original_shape
    n = w * h
    lo, hi = pixels_range
    target = min(hi, max(lo, n))
    factor = (target / n) ** 0.5
    new_w = int(w * factor)
    new_h = int(h * factor)
    new_h = math.ceil(new_h / shape_multiplier) * shape_multiplier
    new_w = math.ceil(new_w / shape_multiplier) * shape_multiplier
    return factor, (new_h, new_w)


def resolve_pixel_bounds(pixels_bounds, resolution_level: Optional[int]):
    """unidepthv2.py:247-262 (resolution_level sub-interval)."""
    if resolution_level is None:
        return tuple(pixels_bounds)
    assert 0 <= resolution_level < 10, "resolution_level should be in [0, 10)"
    interval = (pixels_bounds[1] - pixels_bounds[0]) / 10
    return (resolution_level * interval + pixels_bounds[0],
            (resolution_level + 1) * interval + pixels_bounds[0])


# --------------------------------------------------------------------------------------
# small building blocks
# --------------------------------------------------------------------------------------
def _ln(x, sd, prefix, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def _lin(x, sd, prefix, bias=True):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias") if bias else None)


def _heads(x, h):
    b, n, c = x.shape
    return x.view(b, n, h, c // h).transpose(1, 2)


def _sdpa(q, k, v):
    # softmax(q k^T / sqrt(d)) v, no mask, no dropout (metadinov2/attention.py:58,
    # layers/attention.py:136 call the same torch op)
    return F.scaled_dot_product_attention(q, k, v)


def _mlp(x, sd, prefix, eps=1e-5):
    """layers/mlp.py:9-35: LN -> proj1 -> GELU(erf) -> proj2."""
    x = _ln(x, sd, prefix + ".norm", eps)
    x = F.gelu(_lin(x, sd, prefix + ".proj1"))
    return _lin(x, sd, prefix + ".proj2")


def _attention_block(x, ctx, sd, prefix, heads, pos_q=None):
    """layers/attention.py:81-164 (AttentionBlock): pre-LN cross/self attention with
    optional LayerScale (present iff the state-dict has `<prefix>.ls1.gamma`)."""
    context = x if ctx is None else ctx
    xn = _ln(x, sd, prefix + ".norm_attnx", 1e-5)
    cn = _ln(context, sd, prefix + ".norm_attnctx", 1e-5)
    kv = F.linear(cn, sd[prefix + ".kv.weight"], sd.get(prefix + ".kv.bias"))
    c = x.shape[-1]
    k, v = kv[..., :c], kv[..., c:]            # "(kv h d)": k rows first, then v rows
    q = F.linear(xn, sd[prefix + ".q.weight"], sd.get(prefix + ".q.bias"))
    q, k, v = _heads(q, heads), _heads(k, heads), _heads(v, heads)
    if pos_q is not None:
        q = q + _heads(pos_q, heads)
    o = _sdpa(q, k, v).transpose(1, 2).reshape(x.shape)
    o = F.linear(o, sd[prefix + ".out.weight"], sd.get(prefix + ".out.bias"))
    if prefix + ".ls1.gamma" in sd:
        o = o * sd[prefix + ".ls1.gamma"]
    x = o + x
    m = _mlp(x, sd, prefix + ".mlp")
    if prefix + ".ls2.gamma" in sd:
        m = m * sd[prefix + ".ls2.gamma"]
    return m + x


# --------------------------------------------------------------------------------------
# a2: preprocess
# --------------------------------------------------------------------------------------
def preprocess(rgb: torch.Tensor, paddings, new_hw, normalize=True):
    """unidepthv2.py:288-297: /255, ImageNet standardise, zero-pad AFTER normalisation,
    bilinear (align_corners=False, no antialias) to the network shape."""
    x = rgb.float()
    if normalize:
        mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
        std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
        x = (x / 255.0 - mean) / std
    pl, pr, pt, pb = paddings
    x = F.pad(x, (pl, pr, pt, pb), value=0.0)
    return F.interpolate(x, size=new_hw, mode="bilinear", align_corners=False)


# --------------------------------------------------------------------------------------
# a3-a8: DINOv2 ViT encoder
# --------------------------------------------------------------------------------------
def interpolate_pos_embed(pos_embed: torch.Tensor, gh: int, gw: int):
    """dinov2.py:267-304 with interpolate_offset == 0.0 (V2 factories, encoder.py:190):
    bicubic, antialias=False, size=(gh, gw) from the sqrt(N) x sqrt(N) grid."""
    n = pos_embed.shape[1] - 1
    m = int(math.sqrt(n))
    assert m * m == n
    dim = pos_embed.shape[-1]
    if gh == m and gw == m:
        return pos_embed
    cls_pos = pos_embed[:, :1]
    grid = pos_embed[:, 1:].reshape(1, m, m, dim).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, size=(gh, gw), mode="bicubic", antialias=False)
    grid = grid.permute(0, 2, 3, 1).reshape(1, gh * gw, dim)
    return torch.cat([cls_pos, grid], dim=1)


def vit_encoder(sd: Dict[str, torch.Tensor], spec: ModelSpec, image: torch.Tensor,
                taps_out: Optional[dict] = None):
    """DinoVisionTransformer.forward (dinov2.py:324-347) + tap selection with the "last"
    stacking fn (unidepthv2.py:365-372, utils/misc.py:24).  Returns the 4 tapped feature
    maps [B,gh,gw,D] and cls tokens [B,1,D]."""
    p = "pixel_encoder."
    b, _, hh, ww = image.shape
    gh, gw = hh // PATCH, ww // PATCH
    d, nh = spec.embed_dim, spec.enc_heads
    # patch_embed.py:71-89: conv k=s=14, flatten h-major
    x = F.conv2d(image, sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"],
                 stride=PATCH)
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat([sd[p + "cls_token"].expand(b, -1, -1), x], dim=1)
    x = x + interpolate_pos_embed(sd[p + "pos_embed"].float(), gh, gw)
    if taps_out is not None:
        taps_out["tokens0"] = x.clone()
    feats, clss = [], []
    for i in range(spec.depth):
        bp = f"{p}blocks.{i}."
        # metadinov2/block.py:84-109, attention.py:51-62, mlp.py:35-41; LN eps 1e-6 (dinov2.py:167)
        h1 = _ln(x, sd, bp + "norm1", 1e-6)
        qkv = _lin(h1, sd, bp + "attn.qkv").view(b, -1, 3, nh, d // nh).permute(2, 0, 3, 1, 4)
        a = _sdpa(qkv[0], qkv[1], qkv[2]).transpose(1, 2).reshape(b, -1, d)
        x = x + _lin(a, sd, bp + "attn.proj") * sd[bp + "ls1.gamma"]
        h2 = _ln(x, sd, bp + "norm2", 1e-6)
        h2 = _lin(F.gelu(_lin(h2, sd, bp + "mlp.fc1")), sd, bp + "mlp.fc2")
        x = x + h2 * sd[bp + "ls2.gamma"]
        if taps_out is not None and i == 0:
            taps_out["block0"] = x.clone()
        if (i + 1) in spec.taps:
            # final norm eps 1e-5 (dinov2.py:254 nn.LayerNorm default), applied iff use_norm
            o = _ln(x, sd, p + "norm", 1e-5) if spec.use_norm else x
            clss.append(o[:, :1])
            feats.append(o[:, 1:].reshape(b, gh, gw, d))
    return feats, clss


# --------------------------------------------------------------------------------------
# a9-a17: decoder
# --------------------------------------------------------------------------------------
def camera_head(sd, spec: ModelSpec, cls_tokens: torch.Tensor, net_hw):
    """CameraHead.forward + fill_intrinsics (unidepthv2/decoder.py:85-111)."""
    p = "pixel_decoder.camera_layer."
    t = _mlp(cls_tokens, sd, p + "project")
    pos = sd[p + "latents_pos"].expand(t.shape[0], -1, -1)
    t = _attention_block(t, None, sd, p + "aggregate1", spec.dec_heads, pos_q=pos)
    t = _attention_block(t, None, sd, p + "aggregate2", spec.dec_heads, pos_q=pos)
    x = _mlp(t, sd, p + "out_pinhole").squeeze(-1)          # [B,4]
    hh, ww = net_hw
    diag = (hh ** 2 + ww ** 2) ** 0.5
    corr = torch.tensor([0.7 * diag, 0.7 * diag, ww, hh], dtype=x.dtype)
    vals = torch.stack([x[:, 0].exp(), x[:, 1].exp(), x[:, 2].sigmoid(), x[:, 3].sigmoid()], 1)
    return corr.unsqueeze(0) * vals                          # fx, fy, cx, cy


def rays_from_intrinsics(intr: torch.Tensor, hh: int, ww: int):
    """Decoder.run_camera (unidepthv2/decoder.py:361-403) + coords_grid
    (utils/coordinate.py:4-20): K, analytic K^-1, unit rays at pixel centres. -> K[B,3,3],
    rays [B,H*W,3]."""
    b = intr.shape[0]
    fx, fy, cx, cy = intr.unbind(-1)
    kinv = torch.eye(3).repeat(b, 1, 1)
    kinv[:, 0, 0] = 1.0 / fx
    kinv[:, 1, 1] = 1.0 / fy
    kinv[:, 0, 2] = -cx / fx
    kinv[:, 1, 2] = -cy / fy
    k = torch.eye(3).repeat(b, 1, 1)
    k[:, 0, 0], k[:, 1, 1], k[:, 0, 2], k[:, 1, 2] = fx, fy, cx, cy
    xs = torch.linspace(0.5, ww - 0.5, ww)
    ys = torch.linspace(0.5, hh - 0.5, hh)
    grid = torch.stack([xs.repeat(hh, 1), ys.repeat(ww, 1).t(), torch.ones(hh, ww)], 0).float()
    rays = kinv @ grid.reshape(1, 3, -1).repeat(b, 1, 1)
    rays = rays.reshape(b, 3, hh, ww)
    rays = rays / torch.norm(rays, dim=1, keepdim=True).clamp(min=1e-5)
    return k, rays.flatten(2).transpose(1, 2)


def embed_rays(rays: torch.Tensor, net_hw, grid_hw, hidden: int):
    """DepthHead.embed_rays (unidepthv2/decoder.py:234-253) -> flat_interpolate
    (utils/geometric.py:227-252, antialiased bilinear) -> polar/azimuth ->
    generate_fourier_features (utils/positional_embedding.py:218-256; log bands, sin only)."""
    b = rays.shape[0]
    (hh, ww), (gh, gw) = net_hw, grid_hw
    t = rays.view(b, hh, ww, 3).permute(0, 3, 1, 2)
    t = F.interpolate(t, size=(gh, gw), mode="bilinear", align_corners=False, antialias=True)
    r = t.flatten(2).transpose(1, 2)
    r = r / torch.norm(r, dim=-1, keepdim=True).clip(min=1e-4)
    x, y, z = r[..., 0], r[..., 1], r[..., 2]
    polar = torch.acos(z)
    x_clipped = x.abs().clip(min=1e-3) * (2 * (x >= 0).int() - 1)
    azimuth = torch.atan2(y, x_clipped)
    ang = torch.stack([polar, azimuth], dim=-1)                # [B,N,2]
    bands = hidden // 2
    scales = 2.0 ** torch.linspace(0.0, math.log2(max(gh, gw) // 2), steps=bands)
    return (ang.unsqueeze(-1) * scales * math.pi).sin().flatten(-2)


def _rcu(x, sd, prefix):
    """ResidualConvUnit.forward (layers/upsample.py:171-180): LeakyReLU -> conv3x3 ->
    LeakyReLU -> conv3x3, gamma*out + x (zero padding, no norm)."""
    o = F.leaky_relu(x, 0.01)
    o = F.conv2d(o, sd[prefix + ".conv1.weight"], sd[prefix + ".conv1.bias"], padding=1)
    o = F.leaky_relu(o, 0.01)
    o = F.conv2d(o, sd[prefix + ".conv2.weight"], sd[prefix + ".conv2.bias"], padding=1)
    return sd[prefix + ".gamma"] * o + x


def _reflect_conv3(x, sd, prefix):
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")
    return F.conv2d(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def _head(x, sd, mlp_prefix, lr_prefix, hr_prefix, net_hw):
    """DepthHead.depth_proj / confidence_proj (unidepthv2/decoder.py:284-313) on the finest
    feature map only (the coarser maps are interpolated and discarded by the reference)."""
    t = x.permute(0, 2, 3, 1)
    t = F.layer_norm(t, (t.shape[-1],), sd[mlp_prefix + ".0.weight"], sd[mlp_prefix + ".0.bias"], 1e-5)
    t = F.linear(t, sd[mlp_prefix + ".1.weight"], sd[mlp_prefix + ".1.bias"]).permute(0, 3, 1, 2)
    t = _reflect_conv3(t, sd, lr_prefix)
    t = F.interpolate(t, size=net_hw, mode="bilinear", align_corners=True)
    t = _reflect_conv3(t, sd, hr_prefix + ".0")
    t = F.leaky_relu(t, 0.01)
    return F.conv2d(t, sd[hr_prefix + ".2.weight"], sd[hr_prefix + ".2.bias"])


def decoder(sd, spec: ModelSpec, feats: List[torch.Tensor], clss: List[torch.Tensor], net_hw,
            rays_gt: Optional[torch.Tensor] = None, taps_out: Optional[dict] = None):
    """Decoder.forward (unidepthv2/decoder.py:405-462) with DepthHead.forward (:320-333)."""
    p = "pixel_decoder."
    dl = p + "depth_layer."
    b, gh, gw, _ = feats[0].shape
    hh, ww = net_hw
    n = gh * gw
    # a9 ListAdapter x2 (decoder.py:35-45, :418, :435)
    fl = [F.linear(f.reshape(b, n, -1), sd[f"{p}input_adapter.input_adapters.{i}.weight"],
                   sd[f"{p}input_adapter.input_adapters.{i}.bias"]) for i, f in enumerate(feats)]
    ct = [F.linear(c, sd[f"{p}camera_token_adapter.input_adapters.{i}.weight"],
                   sd[f"{p}camera_token_adapter.input_adapters.{i}.bias"]) for i, c in enumerate(clss)]
    intr = camera_head(sd, spec, torch.cat(ct, dim=1), net_hw)
    kmat, rays = rays_from_intrinsics(intr, hh, ww)
    if rays_gt is not None:
        rays = rays_gt
    remb = embed_rays(rays, net_hw, (gh, gw), spec.hidden)
    if taps_out is not None:
        taps_out["ray_embedding"] = remb.clone()
        taps_out["intrinsics4"] = intr.clone()
    # a13: 4 prompt blocks (decoder.py:255-260), layer_scale=-1 -> no LayerScale, no biases
    cond = [_attention_block(f, remb, sd, f"{dl}prompt_camera.{i}.layers.0", spec.dec_heads)
            for i, f in enumerate(fl)]
    if taps_out is not None:
        taps_out["cond0"] = cond[0].clone()
    # a14 (decoder.py:262-282)
    init_latents = F.linear(cond[0], sd[dl + "to_latents.weight"], sd[dl + "to_latents.bias"])
    to_map = lambda t: t.transpose(1, 2).reshape(b, -1, gh, gw)
    init_latents = to_map(init_latents)
    latents = init_latents
    for i in range(len(spec.dec_depths)):
        k = max(1, 2 * i)
        latents = latents + F.conv_transpose2d(to_map(cond[i + 1]), sd[f"{dl}process_features.{i}.weight"],
                                               sd[f"{dl}process_features.{i}.bias"], stride=k)
        # a15 ResUpsampleBil (layers/upsample.py:183-223)
        for j in range(spec.dec_depths[i]):
            latents = _rcu(latents, sd, f"{dl}ups.{i}.convs.{j}")
        latents = F.conv2d(latents, sd[f"{dl}ups.{i}.up.0.weight"], sd[f"{dl}ups.{i}.up.0.bias"])
        latents = F.interpolate(latents, scale_factor=2, mode="bilinear", align_corners=False)
        if taps_out is not None:
            taps_out[f"ups{i}"] = latents.clone()
    last = len(spec.dec_depths) - 1
    logdepth = _head(latents, sd, f"{dl}depth_mlp.{last}", dl + "to_depth_lr", dl + "to_depth_hr", net_hw)
    logconf = _head(latents, sd, dl + "confidence_mlp", dl + "to_confidence_lr", dl + "to_confidence_hr", net_hw)
    return {
        "radius": torch.exp(logdepth.clip(min=-8.0, max=8.0) + 2.0),
        "confidence": torch.exp(logconf.clip(min=-8.0, max=8.0)),
        "depth_features": init_latents,
        "intrinsics": kmat,
        "rays": rays,
    }


# --------------------------------------------------------------------------------------
# a18: infer
# --------------------------------------------------------------------------------------
def _post(t, shapes, paddings, mode="bilinear"):
    """_postprocess unidepthv2.py:80-89."""
    t = F.interpolate(t, size=shapes, mode=mode, align_corners=False)
    pl, pr, pt, pb = paddings
    return t[..., pt: shapes[0] - pb, pl: shapes[1] - pr]


def pinhole_rays(kmat: torch.Tensor, hh: int, ww: int):
    """GT-camera branch (a19): Pinhole.get_rays == unproject pixel centres with K^-1
    (utils/camera.py:229-273; Camera.get_rays :115-120).  K [B,3,3] -> rays [B,H*W,3]."""
    b = kmat.shape[0]
    xs = torch.linspace(0.5, ww - 0.5, ww)
    ys = torch.linspace(0.5, hh - 0.5, hh)
    grid = torch.stack([xs.repeat(hh, 1), ys.repeat(ww, 1).t(), torch.ones(hh, ww)], 0).float()
    rays = torch.inverse(kmat) @ grid.reshape(1, 3, -1).repeat(b, 1, 1)
    rays = rays / torch.norm(rays, dim=1, keepdim=True).clamp(min=1e-5)
    return rays.transpose(1, 2)


@torch.no_grad()
def infer_v2(sd: Dict[str, torch.Tensor], config: dict, rgb: torch.Tensor,
             resolution_level: Optional[int] = None, normalize: bool = True,
             interpolation_mode: str = "bilinear", taps_out: Optional[dict] = None, camera=None):
    """UniDepthV2.infer (unidepthv2.py:239-339) + encode_decode (:341-379), fp32 on CPU.  `camera` (GT-camera branch,
    :267-303,:361-362; decoder.py:400): a (...,3,3) pinhole K, or a camera object with the reference's
    crop / resize / get_rays methods; its rays replace the predicted ones, the intrinsics output stays the predicted one."""
    spec = ModelSpec(config)
    if rgb.ndim == 3:
        rgb = rgb.unsqueeze(0)
    b, _, h, w = rgb.shape
    bounds = resolve_pixel_bounds(spec.pixels_bounds, resolution_level)
    paddings, (ph, pw) = get_paddings((h, w), spec.ratio_bounds)
    factor, (nh, nw) = get_resize_factor((ph, pw), bounds)
    x = preprocess(rgb, paddings, (nh, nw), normalize)
    if taps_out is not None:
        taps_out["net_input"] = x.clone()
    feats, clss = vit_encoder(sd, spec, x, taps_out)
    if taps_out is not None:
        taps_out["feat3"] = feats[-1].clone()
        taps_out["cls3"] = clss[-1].clone()
    rays_gt = None
    if isinstance(camera, torch.Tensor):
        kn = camera.reshape(-1, 3, 3).float().clone()      # Camera.crop(-pads) then resize(factor): camera.py:78-81,115-120
        kn[:, 0, 2] += paddings[0]
        kn[:, 1, 2] += paddings[2]
        kn[:, :2, :] *= factor
        rays_gt = pinhole_rays(kn.expand(b, 3, 3) if kn.shape[0] == 1 else kn, nh, nw)
    elif camera is not None:
        import copy as _copy
        cam = _copy.deepcopy(camera)
        cam = cam.crop(left=-paddings[0], top=-paddings[2], right=-paddings[1], bottom=-paddings[3])
        cam = cam.resize(factor)
        r = cam.get_rays(shapes=(b, nh, nw))
        rays_gt = r.expand(b, 3, nh, nw).flatten(2).transpose(1, 2)
    out = decoder(sd, spec, feats, clss, (nh, nw), rays_gt=rays_gt, taps_out=taps_out)
    rays_map = out["rays"].transpose(1, 2).reshape(b, 3, nh, nw)
    points = rays_map * out["radius"]
    res = {}
    res["confidence"] = _post(out["confidence"], (ph, pw), paddings, interpolation_mode)
    pts = _post(points, (ph, pw), paddings, interpolation_mode)
    rys = _post(rays_map, (ph, pw), paddings, interpolation_mode)
    # _postprocess_intrinsics unidepthv2.py:92-108
    k = out["intrinsics"].clone()
    k[:, 0, 0] /= factor
    k[:, 1, 1] /= factor
    k[:, 0, 2] /= factor
    k[:, 1, 2] /= factor
    k[:, 0, 2] -= paddings[0]
    k[:, 1, 2] -= paddings[2]
    res["intrinsics"] = k
    res["radius"] = pts.norm(dim=1, keepdim=True)
    res["depth"] = pts[:, -1:]
    res["points"] = pts
    res["rays"] = rys / torch.norm(rys, dim=1, keepdim=True).clip(min=1e-5)
    res["depth_features"] = out["depth_features"]
    return res
