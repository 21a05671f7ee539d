"""Oracle pieces of the UniDepthV1 inference path that need no network (SURVEY.md section 8 row a20, next
row 8f-2): the fixed-shape pre/post-processing arithmetic, ray generation and the (theta, phi, z) -> xyz
conversion.  TEST INFRASTRUCTURE: torch fp32 restatements pinned to the reference's own functions through
tests/golden/v1_parts.npz (oracle/make_golden.py).

Reference: unidepth/models/unidepthv1/unidepthv1.py:30-94 (`_paddings`, `_shapes`, `_preprocess`,
`_postprocess`), unidepth/utils/geometric.py:13-53 (`generate_rays`), :57-73
(`spherical_zbuffer_to_euclidean`)."""
import math
from typing import Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


def v1_shapes(image_hw: Tuple[int, int], network_hw: Tuple[int, int]):
    """Resize so that the image fits inside the fixed network shape keeping its aspect ratio:
    ((h', w'), ratio) with h' = ceil(h*ratio - 0.5) (unidepthv1.py:38-46)."""
    h, w = image_hw
    if network_hw[1] / network_hw[0] > w / h:
        ratio = network_hw[0] / h
    else:
        ratio = network_hw[1] / w
    return (math.ceil(h * ratio - 0.5), math.ceil(w * ratio - 0.5)), ratio


def v1_paddings(resized_hw: Tuple[int, int], network_hw: Tuple[int, int]):
    """Symmetric zero padding up to the network shape -> (left, right, top, bottom) (unidepthv1.py:30-35)."""
    dh, dw = network_hw[0] - resized_hw[0], network_hw[1] - resized_hw[1]
    return dw // 2, dw - dw // 2, dh // 2, dh - dh // 2


def v1_preprocess(rgbs: torch.Tensor, intrinsics: Optional[torch.Tensor], resized_hw, pads, ratio: float):
    """Antialiased bilinear resize, zero pad, and the matching K update (unidepthv1.py:49-63)."""
    left, right, top, bottom = pads
    x = F.interpolate(rgbs, size=tuple(resized_hw), mode="bilinear", align_corners=False, antialias=True)
    x = F.pad(x, (left, right, top, bottom))
    if intrinsics is None:
        return x, None
    k = intrinsics.clone()
    k[:, 0, 0] = k[:, 0, 0] * ratio
    k[:, 1, 1] = k[:, 1, 1] * ratio
    k[:, 0, 2] = k[:, 0, 2] * ratio + left
    k[:, 1, 2] = k[:, 1, 2] * ratio + top
    return x, k


def v1_postprocess(predictions: Sequence[torch.Tensor], intrinsics: torch.Tensor, network_hw, pads, ratio: float,
                   original_hw):
    """Mean of the multi-scale predictions at network resolution, crop the paddings, antialiased resize
    to the original size, and K back to the original frame (unidepthv1.py:66-94)."""
    left, right, top, bottom = pads
    up = [F.interpolate(p, size=tuple(network_hw), mode="bilinear", align_corners=False, antialias=True) for p in predictions]
    pred = sum(up) / len(up)
    pred = pred[..., top:network_hw[0] - bottom, left:network_hw[1] - right]
    pred = F.interpolate(pred, size=tuple(original_hw), mode="bilinear", align_corners=False, antialias=True)
    k = intrinsics.clone()
    k[:, 0, 0] = k[:, 0, 0] / ratio
    k[:, 1, 1] = k[:, 1, 1] / ratio
    k[:, 0, 2] = (k[:, 0, 2] - left) / ratio
    k[:, 1, 2] = (k[:, 1, 2] - top) / ratio
    return pred, k


def generate_rays(intrinsics: torch.Tensor, image_hw: Tuple[int, int]):
    """Unit rays K^-1 [u+0.5, v+0.5, 1] (analytic inverse of a skew-free K) as [B, H*W, 3] and their angles
    (theta = atan2(x, z), phi = acos(y)) as [B, H*W, 2] (geometric.py:13-53)."""
    h, w = image_hw
    fx, fy, cx, cy = intrinsics[:, 0, 0], intrinsics[:, 1, 1], intrinsics[:, 0, 2], intrinsics[:, 1, 2]
    u = torch.arange(w, dtype=intrinsics.dtype) + 0.5
    v = torch.arange(h, dtype=intrinsics.dtype) + 0.5
    x = (u[None, None, :] - cx[:, None, None]) / fx[:, None, None]
    y = (v[None, :, None] - cy[:, None, None]) / fy[:, None, None]
    d = torch.stack([x.expand(-1, h, w), y.expand(-1, h, w), torch.ones(len(fx), h, w, dtype=intrinsics.dtype)], dim=-1)
    d = F.normalize(d.reshape(len(fx), h * w, 3), dim=-1)
    angles = torch.stack([torch.atan2(d[..., 0], d[..., 2]), torch.acos(d[..., 1])], dim=-1)
    return d, angles


def spherical_zbuffer_to_euclidean(tpz: torch.Tensor) -> torch.Tensor:
    """(theta, phi, z) -> (x, y, z) with x = z tan(theta), y = z / tan(phi) / cos(theta) (geometric.py:57-73)."""
    theta, phi, z = tpz[..., 0], tpz[..., 1], tpz[..., 2]
    return torch.stack([z * torch.tan(theta), z / torch.tan(phi) / torch.cos(theta), z], dim=-1)
