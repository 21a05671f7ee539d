"""Random (seeded) weights for benchmarks / smoke runs when no checkpoint is available.

Drawn directly on the target device with a non-degenerate scaling (matmul weights ~ N(0, 1/fan_in),
small biases, LayerScale ~ 0.2) so that depth / confidence / intrinsics are
well spread.  Values are NOT those of oracle/fixture.py (different RNG stream); parity tests that
need both sides to agree copy one state-dict to the other side.
"""
from __future__ import annotations

import torch

from .spec import param_shapes


def synthetic_state_dict(config: dict, seed: int = 0, device="cuda"):
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    for key, shape in param_shapes(config).items():
        n = lambda: torch.randn(shape, generator=g, device=device)
        u = lambda: torch.rand(shape, generator=g, device=device)
        leaf = key.rsplit(".", 1)[-1]
        if key.endswith(("cls_token", "pos_embed")):
            t = 0.2 * n()
        elif key.endswith("latents_pos"):
            t = 0.5 * n()
        elif ".ls1.gamma" in key or ".ls2.gamma" in key:
            t = 0.2 * (0.5 + u())
        elif leaf == "gamma":
            t = 0.5 * (0.5 + u())
        elif "norm" in key or "confidence_mlp.0." in key or (".depth_mlp." in key and key.split(".")[-2] == "0"):
            t = 1.0 + 0.1 * n() if leaf == "weight" else 0.05 * n()
        elif leaf == "bias":
            t = 0.05 * n()
        elif leaf == "weight":
            fan_in = shape[0] if "process_features" in key else int(torch.tensor(shape[1:]).prod())
            t = n() / fan_in ** 0.5
            if key.endswith(("to_depth_hr.2.weight", "to_confidence_hr.2.weight", "out_pinhole.proj2.weight")):
                t = 0.3 * t
        else:
            t = n()
        sd[key] = t.float()
    return sd


def synthetic_state_dict_v1(config: dict, seed: int = 0, device="cuda"):
    """Same idea for UniDepthV1 (ConvNeXt encoder): names / shapes from spec_v1.param_shapes."""
    from .spec_v1 import param_shapes as v1_shapes
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    for key, shape in v1_shapes(config).items():
        n = lambda: torch.randn(shape, generator=g, device=device)
        u = lambda: torch.rand(shape, generator=g, device=device)
        leaf = key.rsplit(".", 1)[-1]
        is_norm = len(shape) == 1 and ("norm" in key or "cls_project.0." in key or "level_embed_layer.3." in key or "stem.1." in key
                                      or "downsample.0." in key or ("input_adapters" in key and key.split(".")[-2] == "0"))
        if key.endswith("mask_token"):
            t = torch.zeros(shape, device=device)
        elif key.endswith(("level_embeds", "latents_pos")):
            t = 0.5 * n()
        elif ".ls1.gamma" in key or ".ls2.gamma" in key:
            t = 0.3 * (0.5 + u())
        elif leaf == "gamma":
            t = 0.4 * (0.5 + u())
        elif is_norm:
            t = 1.0 + 0.1 * n() if leaf == "weight" else 0.05 * n()
        elif leaf == "bias":
            t = 0.05 * n()
        else:
            fan_in = int(torch.tensor(shape[1:]).prod()) if len(shape) > 1 else 1
            t = n() / fan_in ** 0.5
            if key.endswith(("camera_layer.out.proj2.weight", "out2.weight", "out4.weight", "out8.weight")):
                t = 0.3 * t
        sd[key] = t.float()
    return sd
