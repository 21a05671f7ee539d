"""Image-wise data parallel inference: one process per GPU (torchrun), replicated weights, each
rank runs `infer` on its slice of the batch, ONE all-gather collates the outputs
(SURVEY.md section 8e).  The reference has no multi-GPU inference path; this mirrors how its trainer
shards batches over ranks (scripts/train.py:115-136) for the forward pass only."""
from __future__ import annotations

from typing import Dict

import torch
import torch.distributed as dist

_KEYS = ("confidence", "intrinsics", "radius", "depth", "points", "rays", "depth_features")


def _keys(out):
    """Packing order: UniDepthV2's seven outputs, or the subset a model returns (UniDepthV1: intrinsics, depth, points)."""
    return [k for k in _KEYS if k in out]


def shard_bounds(n_images: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of the global batch owned by `rank` (remainder to low ranks)."""
    base, rem = divmod(n_images, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_outputs(out: Dict[str, torch.Tensor]) -> torch.Tensor:
    """Flatten the per-image outputs of one rank into one contiguous f32 buffer [B, F]."""
    b = out["depth"].shape[0]
    return torch.cat([out[k].reshape(b, -1).float() for k in _keys(out)], dim=1).contiguous()


def unpack_outputs(buf: torch.Tensor, like: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    res, off = {}, 0
    n = buf.shape[0]
    for k in _keys(like):
        shp = like[k].shape[1:]
        cnt = 1
        for s in shp:
            cnt *= s
        res[k] = buf[:, off:off + cnt].reshape(n, *shp)
        off += cnt
    return res


class PendingOutputs:
    """Handle of an in-flight all-gather: `.wait()` makes the current stream (NCCL / P2P) or the host
    (gloo) wait for it and returns the global output dict.  Keeps the send / receive buffers alive
    until then."""

    def __init__(self, work, full, local, shapes, event=None):
        self._work, self._full, self._local, self._shapes, self._event = work, full, local, shapes, event

    def __del__(self):
        try:
            if self._event is not None:      # dropped without wait(): still order the free after the copies
                torch.cuda.current_stream().wait_event(self._event)
        except Exception:
            pass

    def wait(self) -> Dict[str, torch.Tensor]:
        if self._work is not None:
            self._work.wait()
            self._work = None
        if self._event is not None:
            torch.cuda.current_stream().wait_event(self._event)
            self._event = None
        return unpack_outputs(self._full, self._shapes)


class P2PGather:
    """All-gather of the packed outputs over NVLink peer memory with the COPY ENGINES: every rank packs
    into a buffer that all peers have mapped (torch symmetric memory), a device-side barrier, then each
    rank pulls the other ranks' buffers with plain device-to-device copies on a side stream.  No SM is
    used by the transfer, which matters here: the GEMM kernels are persistent with one CTA per SM, so
    an overlapping NCCL kernel that occupies even a few SMs stalls whole tile columns (measured at N=2:
    in-flight NCCL all-gather 18.6 ms/step vs 18.0 ms compute alone).  Send buffers rotate (depth 2)."""

    def __init__(self, batch: int, feat: int, device, group=None, depth: int = 2):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.shape = (batch, feat)
        # NOT symm_mem.empty(): that allocates through an implicit torch.cuda.MemPool, and graphs captured
        # afterwards then corrupt the memory of graphs captured before (observed: wrong replays of an
        # older CUDA graph after a new capture); the plain p2p allocation has no such side effect
        self.send = [symm_mem._SymmetricMemory.empty_strided_p2p((batch, feat), (feat, 1), torch.float32,
                                                                 torch.device(device)) for _ in range(depth)]
        self.hdl = [symm_mem.rendezvous(t, self.group) for t in self.send]
        self.stream = torch.cuda.Stream(device=device)
        self.last = [None] * depth      # event after which slot i may be overwritten
        self.i = 0

    def start(self, out: Dict[str, torch.Tensor]) -> PendingOutputs:
        slot = self.i % len(self.send)
        self.i += 1
        cur = torch.cuda.current_stream()
        if self.last[slot] is not None:
            cur.wait_event(self.last[slot])            # every peer has pulled the previous content
        b = out["depth"].shape[0]
        torch.cat([out[k].reshape(b, -1) for k in _keys(out)], dim=1, out=self.send[slot])
        full = torch.empty((self.world * b, self.shape[1]), device=self.send[slot].device, dtype=torch.float32)
        hdl = self.hdl[slot]
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            hdl.barrier()                              # all send buffers of this slot are complete
            for step in range(self.world):
                r = (self.rank - step) % self.world
                full[r * b:(r + 1) * b].copy_(hdl.get_buffer(r, self.shape, torch.float32), non_blocking=True)
            hdl.barrier()                              # all pulls done: the slot may be reused
            ev = torch.cuda.Event()
            ev.record(self.stream)
        # no full.record_stream(): PendingOutputs keeps `full` alive and its wait() orders the consumer
        # stream after the side stream, so the block is never freed with side-stream work pending
        self.last[slot] = ev
        shapes = {k: torch.empty((0,) + tuple(out[k].shape[1:]), device="meta") for k in _keys(out)}
        return PendingOutputs(None, full, None, shapes, event=ev)


def gather_mode() -> str:
    """'nccl' (synchronous all_gather_into_tensor) or 'p2p' (copy-engine pulls over NVLink peer memory)."""
    import os
    return os.environ.get("UDB_GATHER", "nccl")


def gather_description() -> str:
    if _p2p_cache:
        return ("one packed all-gather of the per-rank outputs per step: copy-engine pulls over NVLink peer memory on a side "
                "stream, left in flight under the next step's compute (depth-1 pipeline); every gather completes inside the "
                "timed region")
    return "one packed NCCL all_gather_into_tensor of the per-rank outputs per step, synchronous, inside the timed region"


_p2p_cache: Dict[tuple, "P2PGather"] = {}
_p2p_failed = [None]


def p2p_gather_for(out: Dict[str, torch.Tensor], group=None):
    """Cached P2PGather for this output signature, or None (then the NCCL path is used).
    OPT-IN (UDB_GATHER=p2p), experimental: measured at N=2 it removes the gather from the step time
    (99.0 % weak-scaling efficiency vs 98.4 % for the synchronous NCCL gather and 96.5 % for an in-flight
    NCCL gather), but in tests/test_multigpu_gpu.py's sequence (gather, then capture of a LARGER CUDA
    graph, then replay of the older graph) the older graph's replays return wrong values with this path
    enabled and not with NCCL; not understood yet, so it is off by default."""
    if gather_mode() != "p2p" or _p2p_failed[0] is not None or not out["depth"].is_cuda:
        return None
    b = out["depth"].shape[0]
    feat = sum(out[k][0].numel() for k in _keys(out))
    key = (b, feat, out["depth"].device.index)
    if key not in _p2p_cache:
        try:
            _p2p_cache[key] = P2PGather(b, feat, out["depth"].device, group)
        except Exception as e:      # no NVLink peer access / unsupported build: keep working over NCCL
            _p2p_failed[0] = f"{type(e).__name__}: {e}"
            return None
    return _p2p_cache[key]


def gather_outputs(out: Dict[str, torch.Tensor], world: int, group=None, async_op: bool = False, counts=None):
    """Single all-gather of the packed per-rank buffer.  `counts` = images per rank when they differ
    (uneven tail of the global batch; every rank can derive it from shard_bounds): the local buffer is
    zero-padded to the largest count for the collective and the padding rows are dropped afterwards.
    async_op=True returns a PendingOutputs: the collective then overlaps whatever the caller enqueues
    next (the following micro-batch's infer), which is how a serving loop hides it."""
    if world == 1:
        return PendingOutputs(None, pack_outputs(out), None, out) if async_op else out
    if counts is not None and len(set(counts)) > 1:
        assert len(counts) == world and out["depth"].shape[0] == counts[dist.get_rank(group)]
        bmax = max(counts)
        local = pack_outputs(out)
        if local.shape[0] < bmax:
            local = torch.cat([local, local.new_zeros((bmax - local.shape[0], local.shape[1]))], 0)
        full = torch.empty((world * bmax, local.shape[1]), device=local.device, dtype=local.dtype)
        dist.all_gather_into_tensor(full, local, group=group)
        keep = torch.cat([torch.arange(r * bmax, r * bmax + c) for r, c in enumerate(counts)]).to(full.device)
        res = unpack_outputs(full.index_select(0, keep), out)
        return PendingOutputs(None, pack_outputs(res), None, res) if async_op else res
    if dist.get_backend(group) == "nccl":
        p2p = p2p_gather_for(out, group)
        if p2p is not None:
            pending = p2p.start(out)
            return pending if async_op else pending.wait()
    # depth_features is returned as a permuted view; make the packing layout-independent
    local = pack_outputs(out)
    full = torch.empty((world * local.shape[0], local.shape[1]), device=local.device, dtype=local.dtype)
    work = dist.all_gather_into_tensor(full, local, group=group, async_op=True)
    shapes = {k: torch.empty((0,) + tuple(out[k].shape[1:]), device="meta") for k in _keys(out)}
    pending = PendingOutputs(work, full, local, shapes)
    return pending if async_op else pending.wait()


def infer_sharded(model, rgb: torch.Tensor, async_op: bool = False, **kw):
    """`rgb` is the GLOBAL batch [N,3,H,W] (same on every rank); returns the global outputs on
    every rank (or, with async_op=True, a PendingOutputs whose gather is still in flight).
    N need not be divisible by the world size: low ranks take the remainder (shard_bounds)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    n = rgb.shape[0]
    assert n >= world, "need at least one image per rank"
    lo, hi = shard_bounds(n, rank, world)
    counts = [shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world)]
    kw = dict(kw)
    cam = kw.get("camera", None)
    if cam is not None:
        kw["camera"] = shard_camera(cam, n, lo, hi)
    return gather_outputs(model.infer(rgb[lo:hi], **kw), world, async_op=async_op, counts=counts)


def shard_camera(camera, n: int, lo: int, hi: int):
    """Slice a per-image `camera=` argument the way `rgb` is sliced: a [N,3,3] K tensor or a batched camera
    object (anything indexable whose length is the global batch) gives rank r its images' cameras; a single
    K / camera is shared by every image and passes through."""
    if isinstance(camera, torch.Tensor):
        k = camera.reshape(-1, 3, 3)
        if k.shape[0] == 1:
            return camera
        if k.shape[0] != n:
            raise ValueError(f"camera holds {k.shape[0]} intrinsics for a global batch of {n} images")
        return k[lo:hi]
    try:
        length = len(camera)
    except TypeError:
        return camera
    if length == 1:
        return camera
    if length != n:
        raise ValueError(f"camera holds {length} cameras for a global batch of {n} images")
    return camera[lo:hi]
