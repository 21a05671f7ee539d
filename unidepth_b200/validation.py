"""Host-side glue of the reference's validation path: matching the network
outputs to the ground-truth frame and the scalar depth metrics.  Plain torch ops on small tensors; the
network itself runs in libudb.so (UniDepthV2.forward_test).

Reference: unidepth/utils/misc.py:596-642 (`match_gt`), :645-690 (`match_intrinsics`),
unidepth/utils/evaluation_depth.py:20-34, 93-110 (metrics)."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F


def _pad4(padding, i):
    return tuple(int(v) for v in padding[i]) if padding is not None else (0, 0, 0, 0)


def match_gt(pred: torch.Tensor, target: torch.Tensor, padding1: Optional[Sequence] = None,
             padding2: Optional[Sequence] = None, mode: str = "bilinear") -> torch.Tensor:
    """Per image: strip `padding1` (l, r, t, b) from `pred`, resize to `target`'s un-padded size, re-pad
    with `padding2`; interpolation happens in the target's dtype, the result returns to pred's dtype."""
    out = []
    for i in range(len(pred)):
        l1, r1, t1, b1 = _pad4(padding1, i)
        l2, r2, t2, b2 = _pad4(padding2, i)
        item = pred[i]
        core = item[:, t1:item.shape[1] - b1, l1:item.shape[2] - r1]
        size = (target[i].shape[1] - t2 - b2, target[i].shape[2] - l2 - r2)
        resized = F.interpolate(core.unsqueeze(0).to(target[i].dtype), size=size, mode=mode)
        out.append(F.pad(resized, (l2, r2, t2, b2)))
    return torch.cat(out).to(pred[0].dtype)


def match_intrinsics(K: torch.Tensor, image: torch.Tensor, target: torch.Tensor, padding1: Optional[Sequence] = None,
                     padding2: Optional[Sequence] = None) -> torch.Tensor:
    """Pinhole K of the (padded) network input -> K of the target frame: un-pad, scale each axis by the
    ratio of the un-padded sizes, re-pad."""
    out = K.clone()
    h1, w1 = image.shape[2], image.shape[3]
    h2, w2 = target.shape[2], target.shape[3]
    for i in range(K.shape[0]):
        l1, r1, t1, b1 = _pad4(padding1, i)
        l2, r2, t2, b2 = _pad4(padding2, i)
        sx = (w2 - l2 - r2) / (w1 - l1 - r1)
        sy = (h2 - t2 - b2) / (h1 - t1 - b1)
        out[i, 0, 0] *= sx
        out[i, 1, 1] *= sy
        out[i, 0, 2] = (K[i, 0, 2] - l1) * sx + l2
        out[i, 1, 2] = (K[i, 1, 2] - t1) * sy + t2
    return out


def depth_metrics(gt: torch.Tensor, pred: torch.Tensor, mask: Optional[torch.Tensor] = None) -> Dict[str, float]:
    """Scalar metrics of one image on the valid pixels (the subset of evaluation_depth.py's DICT_METRICS
    that needs no scale-alignment solver): d1/d2/d3, rmse, rmselog, arel, sqrel, log10, silog."""
    if mask is None:
        mask = gt > 0
    g, p = gt[mask].double(), pred[mask].double()
    ratio = torch.maximum(g / p, p / g)
    lg = torch.log(p) - torch.log(g)
    return {
        "d1": (ratio < 1.25).double().mean().item(),
        "d2": (ratio < 1.25 ** 2).double().mean().item(),
        "d3": (ratio < 1.25 ** 3).double().mean().item(),
        "rmse": torch.sqrt(((g - p) ** 2).mean()).item(),
        "rmselog": torch.sqrt((lg ** 2).mean()).item(),
        "arel": ((g - p).abs() / g).mean().item(),
        "sqrel": (((g - p) ** 2) / g).mean().item(),
        "log10": (torch.log10(p) - torch.log10(g)).abs().mean().item(),
        "silog": (100 * torch.std(lg)).item(),
    }
