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


class PendingOutputs:
    """Handle of an in-flight all-gather: `.wait()` makes the current stream (NCCL) / the host (gloo) wait
    for it and returns the global output dict.  Keeps the send / receive buffers alive until then."""

    def __init__(self, work, full, local, shapes):
        self._work, self._full, self._local, self._shapes = work, full, local, shapes

    def wait(self) -> Dict[str, torch.Tensor]:
        if self._work is not None:
            self._work.wait()
            self._work = None
        return unpack_outputs(self._full, self._shapes)


def gather_outputs(out: Dict[str, torch.Tensor], world: int, group=None, async_op: bool = False):
    """Single all-gather of the packed per-rank buffer (equal per-rank batch sizes).
    async_op=True returns a PendingOutputs: the collective then overlaps whatever the caller enqueues
    next (the following micro-batch's infer), which is how a serving loop hides it."""
    if world == 1:
        return PendingOutputs(None, pack_outputs(out), None, out) if async_op else out
    # depth_features is returned as a permuted view; make the packing layout-independent
    local = pack_outputs(out)
    full = torch.empty((world * local.shape[0], local.shape[1]), device=local.device, dtype=local.dtype)
    work = dist.all_gather_into_tensor(full, local, group=group, async_op=True)
    shapes = {k: torch.empty((0,) + tuple(out[k].shape[1:]), device="meta") for k in _KEYS}
    pending = PendingOutputs(work, full, local, shapes)
    return pending if async_op else pending.wait()


def infer_sharded(model, rgb: torch.Tensor, async_op: bool = False, **kw):
    """`rgb` is the GLOBAL batch [N,3,H,W] (same on every rank); returns the global outputs on
    every rank (or, with async_op=True, a PendingOutputs whose gather is still in flight).
    N must be divisible by the world size."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    n = rgb.shape[0]
    assert n % world == 0, "global batch must be divisible by the number of ranks"
    lo, hi = shard_bounds(n, rank, world)
    return gather_outputs(model.infer(rgb[lo:hi], **kw), world, async_op=async_op)
