"""Image-wise data parallel inference: one process per GPU (torchrun), replicated weights, each
rank runs `infer` on its slice of the batch, ONE all-gather collates the outputs
(SURVEY.md section 8e).  The reference has no multi-GPU inference path; this mirrors how its trainer
shards batches over ranks (scripts/train.py:115-136) for the forward pass only."""
from __future__ import annotations

from typing import Dict

import torch
import torch.distributed as dist

_KEYS = ("confidence", "intrinsics", "radius", "depth", "points", "rays", "depth_features")


def shard_bounds(n_images: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of the global batch owned by `rank` (remainder to low ranks)."""
    base, rem = divmod(n_images, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_outputs(out: Dict[str, torch.Tensor]) -> torch.Tensor:
    """Flatten the per-image outputs of one rank into one contiguous f32 buffer [B, F]."""
    b = out["depth"].shape[0]
    return torch.cat([out[k].reshape(b, -1).float() for k in _KEYS], dim=1).contiguous()


def unpack_outputs(buf: torch.Tensor, like: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    res, off = {}, 0
    n = buf.shape[0]
    for k in _KEYS:
        shp = like[k].shape[1:]
        cnt = 1
        for s in shp:
            cnt *= s
        res[k] = buf[:, off:off + cnt].reshape(n, *shp)
        off += cnt
    return res


def gather_outputs(out: Dict[str, torch.Tensor], world: int, group=None) -> Dict[str, torch.Tensor]:
    """Single all-gather of the packed per-rank buffer (equal per-rank batch sizes)."""
    if world == 1:
        return out
    # depth_features is returned as a permuted view; make the packing layout-independent
    local = pack_outputs(out)
    full = torch.empty((world * local.shape[0], local.shape[1]), device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(full, local, group=group)
    return unpack_outputs(full, out)


def infer_sharded(model, rgb: torch.Tensor, **kw) -> Dict[str, torch.Tensor]:
    """`rgb` is the GLOBAL batch [N,3,H,W] (same on every rank); returns the global outputs on
    every rank.  N must be divisible by the world size."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    n = rgb.shape[0]
    assert n % world == 0, "global batch must be divisible by the number of ranks"
    lo, hi = shard_bounds(n, rank, world)
    return gather_outputs(model.infer(rgb[lo:hi], **kw), world)
