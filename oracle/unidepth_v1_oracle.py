"""Torch-fp32 functional restatement of `UniDepthV1.infer()` for the ConvNeXt encoder (BASELINE config 4,
SURVEY.md section 8 rows a20 / f2).  TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's CPU legs, never by the product path.

Pinning.  Everything except the Nystrom attention is pinned to outputs of the UNMODIFIED reference
(tests/golden/v1_*.npz, made by oracle/make_golden.py by importing /root/reference through oracle/ref_shims).
The reference's `NystromBlock` calls `xformers.components.attention.NystromAttention(num_landmarks=128)`
(unidepth/layers/nystrom_attention.py:9,44-46,81; `xformers>=0.0.26`, requirements.txt:24); xformers is not
installed here and is not under /root/reference, and no reference test holds a golden vector for it.  `nystrom_attention`
below restates the PUBLISHED algorithm (Xiong et al. 2021, "Nystromformer", as implemented by xformers' defaults:
segment-mean landmarks, 6 Newton-Schulz iterations for the pseudo-inverse with the exact 1/||K||_1 initialisation, no
convolutional skip connection) per head.  ** PARITY UNPINNED for that one function **: the goldens are generated with
this same function substituted for the missing xformers class, so they pin every other line of the path and the
plumbing around the Nystrom blocks, not xformers' arithmetic.  Second anchor (not a pin to xformers itself):
tests/test_oracle_golden.py::test_nystrom_restatement_matches_independent_implementation checks the function against
the Nystromformer authors' implementation shipped in HuggingFace transformers (present in this image) to 6e-7 on the
decoder's shapes, for sequence lengths that are a multiple of the landmark count.

Reference walk (file:line under /root/reference/unidepth):
  models/unidepthv1/unidepthv1.py:288-373  infer            -> infer_v1
                                  :30-94   _shapes/_paddings/_preprocess/_postprocess -> unidepth_v1_parts
  models/backbones/convnext.py:459-471     ConvNeXt.forward -> convnext_oracle.convnext_encoder
  models/unidepthv1/decoder.py:364-463     Decoder.forward  -> decoder_v1
                                  :311-343 run_camera, :38-106 CameraHead -> camera_head
                                  :195-300 DepthHead.forward -> depth_head
  layers/attention.py:81-164               AttentionBlock   -> attention_block
  layers/nystrom_attention.py:22-84        NystromBlock     -> attention_block(nystrom=True)
  layers/mlp.py:9-35                       MLP              -> mlp
  layers/upsample.py:13-45, convnext.py:5-44  ConvUpsample / CvnxtBlock -> conv_upsample / cvnxt_block
  layers/positional_encoding.py:15-59      PositionEmbeddingSine -> position_embedding_sine
  utils/geometric.py:13-53,57-73,228-252   generate_rays / spherical_zbuffer_to_euclidean / flat_interpolate
  utils/sht.py:833                         rsh_cart_8       -> sh81.rsh_cart
"""
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from convnext_oracle import convnext_encoder
from sh81 import rsh_cart
from unidepth_v1_parts import (generate_rays, spherical_zbuffer_to_euclidean, v1_paddings, v1_postprocess,
                               v1_preprocess, v1_shapes)

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)

CONVNEXT_L = dict(depths=(3, 3, 27, 3), dims=(192, 384, 768, 1536))


# ------------------------------------------------------------------------------------------ layers
def mlp(sd, p: str, x: torch.Tensor) -> torch.Tensor:
    """layers/mlp.py:9-35: LayerNorm -> Linear -> GELU(erf) -> Linear (no residual inside)."""
    x = F.layer_norm(x, (x.shape[-1],), sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-5)
    x = F.gelu(F.linear(x, sd[p + ".proj1.weight"], sd[p + ".proj1.bias"]))
    return F.linear(x, sd[p + ".proj2.weight"], sd[p + ".proj2.bias"])


def _avg_landmarks(x: torch.Tensor, n: int) -> torch.Tensor:
    """[b,h,s,d] -> [b,h,n,d]: mean of n contiguous segments; when s % n != 0 the first n - s % n segments hold
    s // n rows and the last s % n hold one more (xformers' AvgPool)."""
    s = x.shape[-2]
    seg = s // n
    assert seg > 0, "num_landmarks should be smaller than the sequence length"
    if s % n == 0:
        return x.reshape(*x.shape[:-2], n, seg, x.shape[-1]).mean(dim=-2)
    n_round = n - s % n
    a = x[..., : n_round * seg, :].reshape(*x.shape[:-2], n_round, seg, x.shape[-1]).mean(dim=-2)
    b = x[..., n_round * seg:, :].reshape(*x.shape[:-2], n - n_round, seg + 1, x.shape[-1]).mean(dim=-2)
    return torch.cat([a, b], dim=-2)


def _iterative_pinv(k: torch.Tensor, n_iter: int = 6) -> torch.Tensor:
    """Newton-Schulz pseudo-inverse of a row-stochastic matrix (Razavi et al.), Z0 = K^T / ||K||_1."""
    eye = torch.eye(k.shape[-1], dtype=k.dtype, device=k.device)
    v = k.transpose(-1, -2) / k.sum(dim=-2).max(dim=-1).values[..., None, None]
    for _ in range(n_iter):
        kv = k @ v
        v = (0.25 * v) @ (13 * eye - kv @ (15 * eye - kv @ (7 * eye - kv)))
    return v


def nystrom_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, num_landmarks: int = 128) -> torch.Tensor:
    """q,k,v [b,h,s,d] -> [b,h,s,d].  PARITY UNPINNED (module docstring)."""
    d = q.shape[-1]
    if k.shape[-2] == num_landmarks:
        return torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d), -1) @ v
    ql, kl = _avg_landmarks(q, num_landmarks), _avg_landmarks(k, num_landmarks)
    k1 = torch.softmax(q @ kl.transpose(-1, -2) / math.sqrt(d), -1)            # [s, m]
    k2 = torch.softmax(ql @ kl.transpose(-1, -2) / math.sqrt(d), -1)           # [m, m]
    k3 = torch.softmax(ql @ k.transpose(-1, -2) / math.sqrt(d), -1) @ v        # [m, d]
    return (k1 @ _iterative_pinv(k2)) @ k3


def attention_block(sd, p: str, x: torch.Tensor, heads: int, context: Optional[torch.Tensor] = None,
                    pos_embed: Optional[torch.Tensor] = None, pos_embed_context: Optional[torch.Tensor] = None,
                    nystrom: bool = False) -> torch.Tensor:
    """layers/attention.py:110-164 (AttentionBlock) / nystrom_attention.py:48-84: pre-LN cross/self attention with
    positional terms added to q / k after the projections, LayerScale, then the pre-LN MLP."""
    ctx = x if context is None else context
    xn = F.layer_norm(x, (x.shape[-1],), sd[p + ".norm_attnx.weight"], sd[p + ".norm_attnx.bias"], 1e-5)
    cn = F.layer_norm(ctx, (ctx.shape[-1],), sd[p + ".norm_attnctx.weight"], sd[p + ".norm_attnctx.bias"], 1e-5)
    dim = x.shape[-1]
    kv = F.linear(cn, sd[p + ".kv.weight"], sd.get(p + ".kv.bias"))
    q = F.linear(xn, sd[p + ".q.weight"], sd.get(p + ".q.bias"))
    b = x.shape[0]
    split = lambda t: t.reshape(b, t.shape[1], heads, dim // heads).transpose(1, 2)      # b h n d
    k, v = split(kv[..., :dim]), split(kv[..., dim:])
    q = split(q)
    if pos_embed is not None:
        q = q + split(pos_embed)
    if pos_embed_context is not None:
        k = k + split(pos_embed_context)
    if nystrom:
        o = nystrom_attention(q, k, v)
    else:
        o = F.scaled_dot_product_attention(q, k, v)
    o = o.transpose(1, 2).reshape(b, -1, dim)
    o = F.linear(o, sd[p + ".out.weight"], sd.get(p + ".out.bias"))
    x = sd[p + ".ls1.gamma"] * o + x
    return sd[p + ".ls2.gamma"] * mlp(sd, p + ".mlp", x) + x


def cvnxt_block(sd, p: str, x: torch.Tensor) -> torch.Tensor:
    """layers/convnext.py:34-44 on NCHW: depthwise 7x7 (zero pad) -> LN(C, eps 1e-5) -> Linear 4x -> GELU -> Linear -> gamma -> + x."""
    c = x.shape[1]
    y = F.conv2d(x, sd[p + ".dwconv.weight"], sd[p + ".dwconv.bias"], padding=3, groups=c).permute(0, 2, 3, 1)
    y = F.layer_norm(y, (c,), sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-5)
    y = F.linear(F.gelu(F.linear(y, sd[p + ".pwconv1.weight"], sd[p + ".pwconv1.bias"])), sd[p + ".pwconv2.weight"],
                 sd[p + ".pwconv2.bias"])
    return x + (sd[p + ".gamma"] * y).permute(0, 3, 1, 2)


def conv_upsample(sd, p: str, x_flat: torch.Tensor, hw: Tuple[int, int]) -> torch.Tensor:
    """layers/upsample.py:13-45: 2 CvnxtBlocks -> conv1x1 C->C/2 -> UpsamplingBilinear2d(x2) (align_corners=True)
    -> conv3x3 (zero pad); tokens in, tokens out ([b, 4*h*w, C/2])."""
    b, _, c = x_flat.shape
    x = x_flat.transpose(1, 2).reshape(b, c, hw[0], hw[1])
    for j in range(2):
        x = cvnxt_block(sd, f"{p}.convs.{j}", x)
    x = F.conv2d(x, sd[p + ".up.0.weight"], sd[p + ".up.0.bias"])
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    x = F.conv2d(x, sd[p + ".up.2.weight"], sd[p + ".up.2.bias"], padding=1)
    return x.flatten(2).transpose(1, 2)


def position_embedding_sine(h: int, w: int, num_pos_feats: int, dtype=torch.float32) -> torch.Tensor:
    """layers/positional_encoding.py:15-59 with normalize=True, scale 2*pi, temperature 1e4 -> [h*w, 2*num_pos_feats]
    (y features first, then x; sin on even, cos on odd feature indices)."""
    eps, scale = 1e-6, 2 * math.pi
    y = torch.arange(1, h + 1, dtype=torch.float32)[:, None].expand(h, w)
    x = torch.arange(1, w + 1, dtype=torch.float32)[None, :].expand(h, w)
    y = y / (h + eps) * scale
    x = x / (w + eps) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = 10000.0 ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_pos_feats)
    px, py = x[..., None] / dim_t, y[..., None] / dim_t
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=3).flatten(2)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=3).flatten(2)
    return torch.cat((py, px), dim=2).reshape(h * w, 2 * num_pos_feats).to(dtype)


def flat_interpolate(t: torch.Tensor, old: Tuple[int, int], new: Tuple[int, int], antialias: bool = True) -> torch.Tensor:
    """utils/geometric.py:228-252: tokens [b, h*w, c] resampled (bilinear, align_corners=False, antialias) to new grid."""
    if tuple(old) == tuple(new):
        return t
    x = t.reshape(t.shape[0], old[0], old[1], -1).permute(0, 3, 1, 2)
    x = F.interpolate(x, size=tuple(new), mode="bilinear", align_corners=False, antialias=antialias)
    return x.flatten(2).transpose(1, 2).contiguous()


def list_adapter(sd, p: str, xs: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """decoder.py:22-38: per input LayerNorm -> Linear -> GELU."""
    out = []
    for i, x in enumerate(xs):
        a = f"{p}.input_adapters.{i}"
        x = F.layer_norm(x, (x.shape[-1],), sd[a + ".0.weight"], sd[a + ".0.bias"], 1e-5)
        out.append(F.gelu(F.linear(x, sd[a + ".1.weight"], sd[a + ".1.bias"])))
    return out


# ------------------------------------------------------------------------------------------ decoder
def camera_head(sd, p: str, features: List[torch.Tensor], cls_tokens: torch.Tensor, pos_embed: torch.Tensor,
                heads: int) -> torch.Tensor:
    """decoder.py:85-106: -> x [b,4] (log fx, log fy, logit cx, logit cy before exp / sigmoid)."""
    c = F.layer_norm(cls_tokens, (cls_tokens.shape[-1],), sd[p + ".cls_project.0.weight"], sd[p + ".cls_project.0.bias"], 1e-5)
    c = F.linear(F.gelu(F.linear(c, sd[p + ".cls_project.1.weight"], sd[p + ".cls_project.1.bias"])),
                 sd[p + ".cls_project.3.weight"], sd[p + ".cls_project.3.bias"])
    stack = torch.cat(features, dim=1) + pos_embed
    stack = mlp(sd, p + ".in_features", stack)
    ctx = torch.cat((stack, c), dim=1)
    lat = sd[p + ".latents_pos"].expand(c.shape[0], -1, -1)
    c = attention_block(sd, p + ".aggregate", c, 1, context=ctx, pos_embed=lat)
    i = 0
    while f"{p}.layers.{i}.q.weight" in sd:
        c = attention_block(sd, f"{p}.layers.{i}", c, heads, pos_embed=lat)
        i += 1
    return mlp(sd, p + ".out", c).squeeze(-1)


def depth_head(sd, p: str, features: List[torch.Tensor], rays_hr: torch.Tensor, pos_embed: torch.Tensor,
               level_embed: torch.Tensor, shapes: Tuple[int, int], original_hw: Tuple[int, int], heads: int,
               taps: Optional[dict] = None):
    """decoder.py:195-300."""
    h, w = shapes
    emb = {}
    for s, name in ((1, "16"), (2, "8"), (4, "4")):
        r = F.normalize(flat_interpolate(rays_hr, original_hw, (h * s, w * s)), dim=-1)
        emb[name] = mlp(sd, f"{p}.project_rays{name}", rsh_cart(r, 8))
    tokens = torch.cat(features, dim=1)
    tokens_pos = pos_embed + level_embed
    f16_ = F.linear(torch.cat(features, dim=-1), sd[p + ".features_channel_cat.weight"], sd[p + ".features_channel_cat.bias"])
    lat16 = mlp(sd, p + ".to_latents", f16_)
    lat16 = attention_block(sd, p + ".aggregate_16", lat16, 1, context=tokens, pos_embed_context=tokens_pos)
    lat16 = attention_block(sd, p + ".prompt_camera", lat16, 1, context=emb["16"])
    i = 0
    while f"{p}.layers_16.{i}.q.weight" in sd:
        lat16 = attention_block(sd, f"{p}.layers_16.{i}", lat16, heads, pos_embed=emb["16"])
        i += 1
    if taps is not None:
        taps["latents_16"] = lat16
    lat8 = conv_upsample(sd, p + ".up8", lat16 + emb["16"], (h, w))
    conv_out = lambda t, name, hw: F.conv2d(t.transpose(1, 2).reshape(t.shape[0], -1, hw[0], hw[1]),
                                            sd[f"{p}.{name}.weight"], sd[f"{p}.{name}.bias"], padding=1)
    out8 = conv_out(lat8, "out8", (2 * h, 2 * w))
    i = 0
    while f"{p}.layers_8.{i}.q.weight" in sd:
        lat8 = attention_block(sd, f"{p}.layers_8.{i}", lat8, heads // 2, pos_embed=emb["8"], nystrom=True)
        i += 1
    lat4 = conv_upsample(sd, p + ".up4", lat8 + emb["8"], (2 * h, 2 * w))
    out4 = conv_out(lat4, "out4", (4 * h, 4 * w))
    i = 0
    while f"{p}.layers_4.{i}.q.weight" in sd:
        lat4 = attention_block(sd, f"{p}.layers_4.{i}", lat4, heads // 4, pos_embed=emb["4"], nystrom=True)
        i += 1
    lat2 = conv_upsample(sd, p + ".up2", lat4 + emb["4"], (4 * h, 4 * w))
    out2 = conv_out(lat2, "out2", (8 * h, 8 * w))
    outs = [o.clamp(-10.0, 10.0).exp() for o in (out8, out4, out2)]
    return outs, lat16


def decoder_v1(sd, enc_outs: List[torch.Tensor], cls_all: List[torch.Tensor], image_hw: Tuple[int, int],
               output_idx: Sequence[int], heads: int, gt_k: Optional[torch.Tensor] = None, skip_camera: bool = False,
               taps: Optional[dict] = None):
    """decoder.py:345-463 for a pyramid encoder (ConvNeXt): max over each stage's block outputs, common grid =
    second-smallest level, adapters, level / sine position embeddings, camera head, depth head."""
    p = "pixel_decoder."
    H, W = image_hw
    B = enc_outs[0].shape[0]
    ranges = list(zip([0, *output_idx[:-1]], output_idx))
    levels = []
    for i, j in ranges:
        group = enc_outs[i:j]
        levels.append(group[0] if len(group) == 1 else torch.stack(group, dim=-1).max(dim=-1).values)
    n = len(ranges)
    cls_tokens = [cls_all[-i - 1] for i in range(n)]
    resolutions = [tuple(sorted([x.shape[1], x.shape[2]])) for x in levels]
    level_shapes = sorted(set(resolutions))[::-1]
    if len(level_shapes) == 1:
        level_shapes = level_shapes * n
    common = level_shapes[-2]
    flat = [flat_interpolate(x.reshape(B, -1, x.shape[-1]), level_shapes[i], common) for i, x in enumerate(levels)]
    features = list_adapter(sd, p + "input_adapter", flat)
    le = F.linear(F.gelu(F.linear(sd[p + "level_embeds"], sd[p + "level_embed_layer.0.weight"], sd[p + "level_embed_layer.0.bias"])),
                  sd[p + "level_embed_layer.2.weight"], sd[p + "level_embed_layer.2.bias"])
    le = F.layer_norm(le, (le.shape[-1],), sd[p + "level_embed_layer.3.weight"], sd[p + "level_embed_layer.3.bias"], 1e-5)
    hw = common[0] * common[1]
    level_embed = le[:, None, :].expand(n, hw, -1).reshape(1, n * hw, -1).expand(B, -1, -1)
    hidden = le.shape[-1]
    pe = position_embedding_sine(common[0], common[1], hidden // 2).to(le.device)
    pos_embed = pe[None].repeat(B, n, 1)
    if taps is not None:
        taps["features"] = torch.stack(features, dim=-1)
    if not skip_camera:
        toks = list_adapter(sd, p + "token_adapter", cls_tokens)
        x4 = camera_head(sd, p + "camera_layer", features, torch.cat(toks, dim=1), pos_embed + level_embed, heads)
        K = torch.zeros(B, 3, 3, dtype=x4.dtype, device=x4.device)
        K[:, 0, 0] = x4[:, 0].exp() * (max(H, W) / 2)
        K[:, 1, 1] = x4[:, 1].exp() * (max(H, W) / 2)
        K[:, 0, 2] = x4[:, 2].sigmoid() * W
        K[:, 1, 2] = x4[:, 3].sigmoid() * H
        K[:, 2, 2] = 1.0
        rays = generate_rays(K if gt_k is None else gt_k, (H, W))[0]
    else:
        K = gt_k
        rays = generate_rays(gt_k, (H, W))[0]
    outs, lat16 = depth_head(sd, p + "depth_layer", features, rays, pos_embed, level_embed, common, (H, W), heads, taps)
    return K, outs, lat16


# ------------------------------------------------------------------------------------------ infer
@torch.no_grad()
def infer_v1(sd: Dict[str, torch.Tensor], cfg: dict, rgbs: torch.Tensor, intrinsics: Optional[torch.Tensor] = None,
             skip_camera: bool = False, taps: Optional[dict] = None) -> Dict[str, torch.Tensor]:
    """unidepthv1.py:288-373 for the ConvNeXt-L configuration (configs/config_v1_cnvnxtl.json)."""
    if rgbs.ndim == 3:
        rgbs = rgbs.unsqueeze(0)
    if intrinsics is not None and intrinsics.ndim == 2:
        intrinsics = intrinsics.unsqueeze(0)
    B, _, H, W = rgbs.shape
    if rgbs.max() > 5 or rgbs.dtype == torch.uint8:
        rgbs = rgbs.to(torch.float32).div(255)
    if rgbs.min() >= 0.0 and rgbs.max() <= 1.0:
        mean = torch.tensor(IMAGENET_MEAN, dtype=rgbs.dtype).view(1, 3, 1, 1)
        std = torch.tensor(IMAGENET_STD, dtype=rgbs.dtype).view(1, 3, 1, 1)
        rgbs = (rgbs - mean) / std
    net_hw = tuple(cfg["data"]["image_shape"])
    arch = cfg["model"]["pixel_encoder"].get("arch", CONVNEXT_L)
    output_idx = cfg["model"]["pixel_encoder"].get("output_idx", [3, 6, 33, 36])
    (h, w), ratio = v1_shapes((H, W), net_hw)
    pads = v1_paddings((h, w), net_hw)
    x, gt_k = v1_preprocess(rgbs, intrinsics, (h, w), pads, ratio)
    enc_outs, cls_all = convnext_encoder(sd, x, arch["depths"], prefix="pixel_encoder.")
    if taps is not None:
        taps["enc_last"] = enc_outs[-1]
    K, outs, lat16 = decoder_v1(sd, enc_outs, cls_all, net_hw, output_idx, cfg["model"]["num_heads"], gt_k=gt_k,
                                skip_camera=skip_camera and gt_k is not None, taps=taps)
    if taps is not None:
        taps["out8"], taps["out4"], taps["out2"], taps["K_net"] = outs[0], outs[1], outs[2], K.clone()
    pred, K_out = v1_postprocess(outs, K.clone(), net_hw, pads, ratio, (H, W))
    # unidepthv1.py:354-356: with GT intrinsics the reference back-projects with the PRE-PROCESSED K (scaled by `ratio`,
    # shifted by the paddings) on the original-resolution pixel grid -- reproduced as is
    use_k = gt_k if gt_k is not None else K_out
    angles = generate_rays(use_k, (H, W))[1]
    angles = angles.transpose(1, 2).reshape(B, 2, H, W)
    pts = spherical_zbuffer_to_euclidean(torch.cat((angles, pred), dim=1).permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
    return {"intrinsics": K_out, "points": pts, "depth": pred[:, -1:]}
