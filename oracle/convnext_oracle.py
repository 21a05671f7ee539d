"""Functional torch-fp32 restatement of the reference's ConvNeXt encoder forward (UniDepthV1's
`pixel_encoder` for the cnvnxtl config).  TEST INFRASTRUCTURE; pinned to the
reference's own module (run through the timm stand-ins of oracle/ref_shims) in tests/golden/convnext_small.npz.

Reference: unidepth/models/backbones/convnext.py -- ConvNeXt.forward :459-471, ConvNeXtStage.forward :289-298
(downsample = LayerNorm2d + conv k2 s2, :252-263), ConvNeXtBlock.forward :208-223 (depthwise 7x7 -> LayerNorm
(channels last, eps 1e-6) -> Linear 4x -> GELU(erf) -> Linear -> gamma -> + shortcut), patch stem :371-383."""
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F


def _ln_channels(x_nchw: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    return F.layer_norm(x_nchw.permute(0, 2, 3, 1), (x_nchw.shape[1],), w, b, eps).permute(0, 3, 1, 2)


def convnext_encoder(sd: Dict[str, torch.Tensor], x: torch.Tensor, depths: Sequence[int], prefix: str = "",
                     kernel_size: int = 7) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
    """x [B,3,H,W] (normalised) -> (one NHWC feature map per block, one [B,1,C] mean token per block)."""
    g = lambda k: sd[prefix + k]
    patch = g("stem.0.weight").shape[-1]
    x = F.conv2d(x, g("stem.0.weight"), g("stem.0.bias"), stride=patch)
    x = _ln_channels(x, g("stem.1.weight"), g("stem.1.bias"))
    outs = []
    for i, depth in enumerate(depths):
        s = f"stages.{i}."
        if (prefix + s + "downsample.1.weight") in sd:
            x = _ln_channels(x, g(s + "downsample.0.weight"), g(s + "downsample.0.bias"))
            x = F.conv2d(x, g(s + "downsample.1.weight"), g(s + "downsample.1.bias"), stride=2)
        for j in range(depth):
            b = f"{s}blocks.{j}."
            c = x.shape[1]
            y = F.conv2d(x, g(b + "conv_dw.weight"), g(b + "conv_dw.bias"), padding=kernel_size // 2, groups=c)
            y = F.layer_norm(y.permute(0, 2, 3, 1), (c,), g(b + "norm.weight"), g(b + "norm.bias"), 1e-6)
            y = F.linear(F.gelu(F.linear(y, g(b + "mlp.fc1.weight"), g(b + "mlp.fc1.bias"))), g(b + "mlp.fc2.weight"),
                         g(b + "mlp.fc2.bias"))
            x = x + (y * g(b + "gamma")).permute(0, 3, 1, 2)
            outs.append(x.permute(0, 2, 3, 1).contiguous())
    return outs, [o.mean(dim=(1, 2)).unsqueeze(1) for o in outs]
