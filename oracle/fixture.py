"""Seeded synthetic weights shared by the oracle, the reference (when generating golden
vectors) and the CUDA path  --  TEST INFRASTRUCTURE ONLY (see oracle/unidepth_oracle.py).

No pretrained UniDepth/DINOv2 weights exist offline, and the reference's default init
(trunc-normal std 0.02, zero bias, gamma 1) yields near-constant outputs that hide numerical
errors.  `make_state_dict` therefore draws an "amplified" fixture: every
matmul weight ~ N(0, 1/fan_in), non-trivial biases / LayerNorm affine / LayerScale, so depth,
confidence and intrinsics are well spread and nothing sits on the +-8 log clip.

`param_shapes` enumerates the reference's state-dict (key -> shape) from the config alone;
oracle/make_golden.py asserts that it equals the live reference `state_dict()` key-for-key.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import torch

from unidepth_oracle import ModelSpec


def param_shapes(config: dict) -> "OrderedDict[str, tuple]":
    s = ModelSpec(config)
    d, h = s.embed_dim, s.hidden
    out: "OrderedDict[str, tuple]" = OrderedDict()
    p = "pixel_encoder."
    out[p + "cls_token"] = (1, 1, d)
    out[p + "pos_embed"] = (1, 1 + 37 * 37, d)
    out[p + "register_tokens"] = (1, 1, d)
    out[p + "mask_token"] = (1, d)
    out[p + "patch_embed.proj.weight"] = (d, 3, 14, 14)
    out[p + "patch_embed.proj.bias"] = (d,)
    for i in range(s.depth):
        b = f"{p}blocks.{i}."
        out[b + "norm1.weight"] = (d,)
        out[b + "norm1.bias"] = (d,)
        out[b + "attn.qkv.weight"] = (3 * d, d)
        out[b + "attn.qkv.bias"] = (3 * d,)
        out[b + "attn.proj.weight"] = (d, d)
        out[b + "attn.proj.bias"] = (d,)
        out[b + "ls1.gamma"] = (d,)
        out[b + "norm2.weight"] = (d,)
        out[b + "norm2.bias"] = (d,)
        out[b + "mlp.fc1.weight"] = (4 * d, d)
        out[b + "mlp.fc1.bias"] = (4 * d,)
        out[b + "mlp.fc2.weight"] = (d, 4 * d)
        out[b + "mlp.fc2.bias"] = (d,)
        out[b + "ls2.gamma"] = (d,)
    out[p + "norm.weight"] = (d,)
    out[p + "norm.bias"] = (d,)

    p = "pixel_decoder."
    out[p + "level_embeds"] = (1, 1, 4, h)
    for name in ("input_adapter", "camera_token_adapter"):
        for i in range(4):
            out[f"{p}{name}.input_adapters.{i}.weight"] = (h, d)
            out[f"{p}{name}.input_adapters.{i}.bias"] = (h,)

    def mlp(prefix, hid, outd):
        out[prefix + ".norm.weight"] = (h,)
        out[prefix + ".norm.bias"] = (h,)
        out[prefix + ".proj1.weight"] = (hid, h)
        out[prefix + ".proj1.bias"] = (hid,)
        out[prefix + ".proj2.weight"] = (outd, hid)
        out[prefix + ".proj2.bias"] = (outd,)

    def attn_block(prefix, layer_scale):
        mlp(prefix + ".mlp", s.expansion * h, h)
        out[prefix + ".kv.weight"] = (2 * h, h)
        out[prefix + ".q.weight"] = (h, h)
        out[prefix + ".norm_attnx.weight"] = (h,)
        out[prefix + ".norm_attnx.bias"] = (h,)
        out[prefix + ".norm_attnctx.weight"] = (h,)
        out[prefix + ".norm_attnctx.bias"] = (h,)
        out[prefix + ".out.weight"] = (h, h)
        if layer_scale:
            out[prefix + ".ls1.gamma"] = (h,)
            out[prefix + ".ls2.gamma"] = (h,)

    c = p + "camera_layer."
    out[c + "latents_pos"] = (1, 4, h)
    attn_block(c + "aggregate1", True)
    attn_block(c + "aggregate2", True)
    mlp(c + "project", h, h)
    mlp(c + "out_pinhole", h, 1)

    dl = p + "depth_layer."
    n_up = len(s.dec_depths)
    cur, nxt, outd = [], [], []
    for i in range(n_up):
        cur.append(min(h, 2 * h // int(2 ** i)))
        nxt.append(2 * h // int(2 ** (i + 1)))
        outd.append(max(nxt[-1], s.out_dim))
    ks = s.kernel_size
    for i in range(n_up):
        for j in range(s.dec_depths[i]):
            u = f"{dl}ups.{i}.convs.{j}."
            out[u + "gamma"] = (1, cur[i], 1, 1)
            out[u + "conv1.weight"] = (cur[i], cur[i], ks, ks)
            out[u + "conv1.bias"] = (cur[i],)
            out[u + "conv2.weight"] = (cur[i], cur[i], ks, ks)
            out[u + "conv2.bias"] = (cur[i],)
        out[f"{dl}ups.{i}.up.0.weight"] = (outd[i], cur[i], 1, 1)
        out[f"{dl}ups.{i}.up.0.bias"] = (outd[i],)
    last = n_up - 1
    out[f"{dl}depth_mlp.{last}.0.weight"] = (nxt[last],)
    out[f"{dl}depth_mlp.{last}.0.bias"] = (nxt[last],)
    out[f"{dl}depth_mlp.{last}.1.weight"] = (outd[last], nxt[last])
    out[f"{dl}depth_mlp.{last}.1.bias"] = (outd[last],)
    for i in range(n_up):
        k = max(1, 2 * i)
        out[f"{dl}process_features.{i}.weight"] = (h, cur[i], k, k)
        out[f"{dl}process_features.{i}.bias"] = (cur[i],)
    for i in range(4):
        attn_block(f"{dl}prompt_camera.{i}.layers.0", False)
    out[dl + "to_latents.weight"] = (h, h)
    out[dl + "to_latents.bias"] = (h,)
    out[dl + "confidence_mlp.0.weight"] = (nxt[last],)
    out[dl + "confidence_mlp.0.bias"] = (nxt[last],)
    out[dl + "confidence_mlp.1.weight"] = (outd[last], nxt[last])
    out[dl + "confidence_mlp.1.bias"] = (outd[last],)
    od = outd[last]
    for nm in ("to_depth_lr", "to_confidence_lr"):
        out[f"{dl}{nm}.weight"] = (od // 2, od, 3, 3)
        out[f"{dl}{nm}.bias"] = (od // 2,)
    for nm in ("to_depth_hr", "to_confidence_hr"):
        out[f"{dl}{nm}.0.weight"] = (32, od // 2, 3, 3)
        out[f"{dl}{nm}.0.bias"] = (32,)
        out[f"{dl}{nm}.2.weight"] = (1, 32, 1, 1)
        out[f"{dl}{nm}.2.bias"] = (1,)
    return out


def make_state_dict(config: dict, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Amplified seeded fixture (fp32, CPU).  One torch.Generator per tensor, seeded from
    (seed, index in param_shapes), so a tensor's values do not depend on the others."""
    shapes = param_shapes(config)
    sd: Dict[str, torch.Tensor] = {}
    for idx, (key, shape) in enumerate(shapes.items()):
        g = torch.Generator().manual_seed(seed * 1_000_003 + idx)
        n = lambda *s: torch.randn(*s, generator=g)
        u = lambda *s: torch.rand(*s, generator=g)
        leaf = key.rsplit(".", 1)[-1]
        if key.endswith(("register_tokens", "mask_token", "level_embeds")):
            t = n(*shape)                      # dead parameters on the infer path
        elif key.endswith(("cls_token", "pos_embed")):
            t = 0.2 * n(*shape)
        elif key.endswith("latents_pos"):
            t = 0.5 * n(*shape)
        elif ".ls1.gamma" in key or ".ls2.gamma" in key:
            t = 0.2 * (0.5 + u(*shape))
        elif leaf == "gamma":                  # RCU gamma [1,C,1,1]
            t = 0.5 * (0.5 + u(*shape))
        elif ("norm" in key or "confidence_mlp.0." in key
              or (".depth_mlp." in key and key.split(".")[-2] == "0")):
            t = 1.0 + 0.1 * n(*shape) if leaf == "weight" else 0.05 * n(*shape)
        elif leaf == "bias":
            t = 0.05 * n(*shape)
        elif leaf == "weight":
            if "process_features" in key:      # ConvTranspose2d [Cin,Cout,k,k]: fan_in = Cin
                fan_in = shape[0]
            else:
                fan_in = 1
                for s_ in shape[1:]:
                    fan_in *= s_
            t = n(*shape) / fan_in ** 0.5
            if key.endswith(("to_depth_hr.2.weight", "to_confidence_hr.2.weight",
                             "out_pinhole.proj2.weight")):
                t = 0.3 * t
        else:
            raise KeyError(key)
        sd[key] = t.float().contiguous()
    return sd


def convnext_param_shapes(depths, dims, patch=4, kernel_size=7, mlp_ratio=4):
    """key -> shape of the reference ConvNeXt encoder's state_dict (backbones/convnext.py:301-448)."""
    from collections import OrderedDict
    out = OrderedDict()
    out["mask_token"] = (1, dims[0], 1, 1)
    out["stem.0.weight"], out["stem.0.bias"] = (dims[0], 3, patch, patch), (dims[0],)
    out["stem.1.weight"], out["stem.1.bias"] = (dims[0],), (dims[0],)
    prev = dims[0]
    for i, (depth, c) in enumerate(zip(depths, dims)):
        s = f"stages.{i}."
        if i > 0:
            out[s + "downsample.0.weight"], out[s + "downsample.0.bias"] = (prev,), (prev,)
            out[s + "downsample.1.weight"], out[s + "downsample.1.bias"] = (c, prev, 2, 2), (c,)
        for j in range(depth):
            b = f"{s}blocks.{j}."
            out[b + "gamma"] = (c,)
            out[b + "conv_dw.weight"], out[b + "conv_dw.bias"] = (c, 1, kernel_size, kernel_size), (c,)
            out[b + "norm.weight"], out[b + "norm.bias"] = (c,), (c,)
            out[b + "mlp.fc1.weight"], out[b + "mlp.fc1.bias"] = (mlp_ratio * c, c), (mlp_ratio * c,)
            out[b + "mlp.fc2.weight"], out[b + "mlp.fc2.bias"] = (c, mlp_ratio * c), (c,)
        prev = c
    return out


def make_convnext_state_dict(depths, dims, seed: int):
    """Seeded, well-conditioned weights for the ConvNeXt encoder oracle (same recipe idea as make_state_dict)."""
    sd = {}
    for idx, (k, shp) in enumerate(convnext_param_shapes(depths, dims).items()):
        g = torch.Generator().manual_seed(seed * 1_000_003 + idx)
        if k.endswith("gamma"):
            t = 0.5 + torch.rand(shp, generator=g)
        elif len(shp) == 1 and k.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif len(shp) == 1:
            t = 0.05 * torch.randn(shp, generator=g)
        elif k == "mask_token":
            t = torch.zeros(shp)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            t = torch.randn(shp, generator=g) / fan_in ** 0.5
        sd[k] = t
    return sd


def make_v1_state_dict(config: dict, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded, well-conditioned fixture for UniDepthV1 (ConvNeXt encoder): names / shapes from
    unidepth_b200.spec_v1.param_shapes (oracle/make_golden_v1.py asserts they equal the reference model's)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from unidepth_b200.spec_v1 import param_shapes as v1_shapes
    sd: Dict[str, torch.Tensor] = {}
    for idx, (key, shape) in enumerate(v1_shapes(config).items()):
        g = torch.Generator().manual_seed(seed * 1_000_003 + idx)
        n = lambda *s: torch.randn(*s, generator=g)
        u = lambda *s: torch.rand(*s, generator=g)
        leaf = key.rsplit(".", 1)[-1]
        is_norm = ("norm" in key or key.endswith((".0.weight", ".0.bias")) and "input_adapters" in key
                   or "cls_project.0." in key or "level_embed_layer.3." in key or "stem.1." in key
                   or "downsample.0." in key)
        if key.endswith("mask_token"):
            t = torch.zeros(*shape)
        elif key.endswith("level_embeds"):
            t = 0.5 * n(*shape)
        elif key.endswith("latents_pos"):
            t = 0.5 * n(*shape)
        elif ".ls1.gamma" in key or ".ls2.gamma" in key:
            t = 0.3 * (0.5 + u(*shape))
        elif leaf == "gamma":
            t = 0.4 * (0.5 + u(*shape))
        elif is_norm and len(shape) == 1:
            t = 1.0 + 0.1 * n(*shape) if leaf == "weight" else 0.05 * n(*shape)
        elif leaf == "bias":
            t = 0.05 * n(*shape)
        elif leaf == "weight":
            fan_in = 1
            for s_ in shape[1:]:
                fan_in *= s_
            t = n(*shape) / fan_in ** 0.5
            if key.endswith(("camera_layer.out.proj2.weight", "out2.weight", "out4.weight", "out8.weight")):
                t = 0.3 * t
        else:
            raise KeyError(key)
        sd[key] = t.float().contiguous()
    return sd
