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
        if isinstance(self._full, dict):      # PeerGather: already one tensor per key
            return self._full
        return unpack_outputs(self._full, self._shapes)


class _DevView:
    """Exposes a raw device pointer to torch through the CUDA array interface (no copy, no ownership)."""

    def __init__(self, ptr: int, shape, typestr: str = "<f4"):
        self.__cuda_array_interface__ = dict(shape=tuple(int(s) for s in shape), typestr=typestr, data=(int(ptr), False), version=3,
                                             strides=None)


class PeerGather:
    """All-gather of the per-rank outputs over NVLink peer memory with the COPY ENGINES (include/udb.h `udb_p2p_*`):
    every rank owns `depth` send slots in a cudaMalloc'ed buffer whose CUDA-IPC handle its peers have opened; a step is
    (device-side barrier: all slots written) -> each rank pulls every peer's slot, key by key, straight into the final
    per-key output tensors [world*B, ...] with plain device-to-device copies on a side stream -> (barrier: slot reusable).
    No SM moves data and nothing is packed or unpacked: `views()` hands out tensors that alias the next slot, so `infer`
    can write its outputs directly into it (UniDepthV2.output_buffers).  This matters here because the GEMM kernels are
    persistent with one CTA per SM: a collective kernel that occupies even a few SMs stalls whole tile columns (measured:
    in-flight NCCL all-gather 18.6 ms/step vs 18.0 ms compute alone at N=2)."""

    def __init__(self, out_like: Dict[str, torch.Tensor], device, group=None, depth: int = 2):
        import ctypes as C
        from . import _cabi as cabi
        self.cabi, self.C = cabi, C
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.device = torch.device(device)
        self.keys = _keys(out_like)
        self.shapes = {k: tuple(out_like[k].shape) for k in self.keys}
        self.batch = self.shapes[self.keys[0]][0]
        # Physical layout per key: a permuted view of a contiguous tensor (UniDepthV2's depth_features is [B,h,w,C] memory
        # viewed as [B,C,h,w]) keeps its memory order in the slot and in the gathered tensor, so filling the slot is a
        # plain copy instead of a transposing one.  order = dims by decreasing stride (batch first), inv = its inverse.
        self.order, self.inv = {}, {}
        for k in self.keys:
            t = out_like[k]
            order = sorted(range(t.ndim), key=lambda d: (-t.stride(d), d))
            if order[0] != 0 or t.permute(order).is_contiguous() is False:
                order = list(range(t.ndim))
            self.order[k], self.inv[k] = order, [order.index(d) for d in range(t.ndim)]
        self.offs, off = {}, 0
        for k in self.keys:
            self.offs[k] = off
            n = 1
            for d in self.shapes[k]:
                n *= d
            off += (n * 4 + 255) // 256 * 256          # bytes, every key 256-byte aligned inside a slot
        self.slot_bytes = off
        self.depth = depth
        self.flags_off = depth * self.slot_bytes
        total = self.flags_off + 1024                   # [world] uint32 flags + a timeout word at +512
        lib = cabi.lib()
        flag_dev = self.device

        def agree(ok: bool, what: str):
            """Every fallible LOCAL step is followed by a collective vote, so that either all ranks continue or all give up
            (a rank that raised alone would leave its peers waiting in the next collective)."""
            t = torch.tensor([1 if ok else 0], device=flag_dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
            if int(t.item()) == 0:
                raise RuntimeError(f"peer-memory gather unavailable: {what} failed on at least one rank"
                                   + ("" if ok else f" (here: {lib.udb_last_error().decode()})"))

        base, handle = C.c_void_p(), (C.c_char * 64)()
        with torch.cuda.device(self.device):
            rc = lib.udb_p2p_alloc(total, C.byref(base), handle)
        agree(rc == 0, "udb_p2p_alloc")
        self.base = base.value
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=self.group)
        self.peer, ok = [], True
        for r, h in enumerate(handles):
            if r == self.rank:
                self.peer.append(self.base)
                continue
            p = C.c_void_p()
            buf = (C.c_char * 64).from_buffer_copy(h)
            with torch.cuda.device(self.device):
                ok = ok and lib.udb_p2p_open(buf, C.byref(p)) == 0
            self.peer.append(p.value or 0)
        agree(ok, "udb_p2p_open")
        self.peer_flags = torch.tensor([p + self.flags_off for p in self.peer], dtype=torch.int64, device=self.device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.last = [None] * depth                      # event after which slot i may be overwritten
        self.i = 0
        self.epoch = 0
        phys = lambda k, b: [b] + [self.shapes[k][d] for d in self.order[k][1:]]
        self._phys = phys
        self._views = [{k: torch.as_tensor(_DevView(self.base + s * self.slot_bytes + self.offs[k], phys(k, self.batch)),
                                           device=self.device).permute(self.inv[k]) for k in self.keys} for s in range(depth)]
        dist.barrier(group=self.group)                  # every rank has mapped every peer before the first device barrier

    def views(self) -> Dict[str, torch.Tensor]:
        """Tensors aliasing the slot the NEXT start() sends (write the rank's outputs here to skip the staging copy).
        The caller's stream is made to wait until every peer has pulled the slot's previous content."""
        slot = self.i % self.depth
        if self.last[slot] is not None:
            torch.cuda.current_stream().wait_event(self.last[slot])
        return self._views[slot]

    def _barrier(self):
        self.epoch += 1
        C, lib = self.C, self.cabi.lib()
        self.cabi.check(lib.udb_p2p_barrier(C.c_void_p(self.peer_flags.data_ptr()), C.c_void_p(self.base + self.flags_off), self.rank, self.world,
                                            self.epoch, C.c_void_p(self.base + self.flags_off + 512), C.c_void_p(self.stream.cuda_stream)),
                        "udb_p2p_barrier")

    def start(self, out: Dict[str, torch.Tensor]) -> "PendingOutputs":
        slot = self.i % self.depth
        views = self.views()
        self.i += 1
        cur = torch.cuda.current_stream()
        for k in self.keys:
            if out[k].data_ptr() != views[k].data_ptr():
                views[k].copy_(out[k])                  # outputs were not produced in place: one staging copy
        b = self.batch
        full = {k: torch.empty(self._phys(k, self.world * b), device=self.device, dtype=torch.float32).permute(self.inv[k])
                for k in self.keys}
        C, lib = self.C, self.cabi.lib()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            self._barrier()                             # all send slots of this step are complete
            import os
            self_on_main = os.environ.get("UDB_P2P_SELFCOPY", "dma") == "main"      # experiment: own part copied by a kernel
            for step in range(self.world):
                r = (self.rank - step) % self.world     # start with the local slot, then walk the ring
                if r == self.rank and self_on_main:
                    continue
                for k in self.keys:
                    n = full[k][0].numel() * b * 4          # batch is the outermost physical dim: rank r's rows are contiguous
                    dst = full[k].data_ptr() + r * n
                    self.cabi.check(lib.udb_p2p_copy(C.c_void_p(dst), C.c_void_p(self.peer[r] + slot * self.slot_bytes + self.offs[k]), n,
                                                     C.c_void_p(self.stream.cuda_stream)), "udb_p2p_copy")
            self._barrier()                             # all pulls done: the slot may be reused
            ev = torch.cuda.Event()
            ev.record(self.stream)
        if self_on_main:
            for k in self.keys:
                full[k][self.rank * b:(self.rank + 1) * b].copy_(views[k])
        # (results keep each key's memory order: depth_features comes back as the same kind of permuted view infer returns)
        self.last[slot] = ev
        return PendingOutputs(None, full, None, None, event=ev)

    def close(self):
        """Unmap the peers' buffers and free this rank's (after a synchronise; collective order is the caller's business)."""
        if getattr(self, "base", None) is None:
            return
        torch.cuda.synchronize(self.device)
        lib = self.cabi.lib()
        for r, p in enumerate(self.peer):
            if r != self.rank and p:
                lib.udb_p2p_close(self.C.c_void_p(p))
        lib.udb_p2p_free(self.C.c_void_p(self.base))
        self.base, self._views = None, []

    def timed_out(self) -> bool:
        """True if a device barrier gave up waiting for a peer (synchronises)."""
        torch.cuda.synchronize(self.device)
        flag = torch.as_tensor(_DevView(self.base + self.flags_off + 512, (1,), "<i4"), device=self.device)
        return bool(flag.item())


def gather_mode() -> str:
    """'p2p' (default on NCCL process groups: copy-engine pulls over NVLink peer memory) or 'nccl' (UDB_GATHER=nccl:
    one synchronous all_gather_into_tensor of the packed buffer)."""
    import os
    return os.environ.get("UDB_GATHER", "p2p")


def gather_description() -> str:
    if _p2p_cache:
        return ("all-gather of the per-rank outputs by copy-engine pulls over NVLink peer memory (CUDA IPC buffers, device-side "
                "flag barrier, include/udb.h udb_p2p_*), on a side stream, left in flight under the next step's compute (depth-1 "
                "pipeline); outputs are written straight into the send slot (no pack / unpack); every gather completes inside "
                "the timed region")
    return "one packed NCCL all_gather_into_tensor of the per-rank outputs per step, synchronous, inside the timed region"


_p2p_cache: Dict[tuple, "PeerGather"] = {}
_p2p_failed = [None]


def p2p_gather_for(out: Dict[str, torch.Tensor], group=None):
    """Cached PeerGather for this output signature, or None (then the NCCL path is used: UDB_GATHER=nccl, CPU tensors,
    or peer memory unavailable -- the reason is kept in _p2p_failed[0])."""
    if gather_mode() != "p2p" or _p2p_failed[0] is not None or not out["depth"].is_cuda:
        return None
    key = tuple((k, tuple(out[k].shape)) for k in _keys(out)) + (out["depth"].device.index,)
    if key not in _p2p_cache:
        try:
            _p2p_cache[key] = PeerGather(out, out["depth"].device, group)
        except Exception as e:      # agreed on by all ranks (PeerGather votes after every local step): fall back to NCCL
            _p2p_failed[0] = f"{type(e).__name__}: {e}"
            return None
    return _p2p_cache[key]


def output_views(out_like: Dict[str, torch.Tensor]):
    """The send-slot tensors of the PeerGather that serves outputs shaped like `out_like`, or None if there is none (yet):
    assign them to `model.output_buffers` before `infer` so that the outputs land in the slot without a staging copy."""
    if not _p2p_cache or not out_like["depth"].is_cuda:
        return None
    key = tuple((k, tuple(out_like[k].shape)) for k in _keys(out_like)) + (out_like["depth"].device.index,)
    pg = _p2p_cache.get(key)
    return pg.views() if pg is not None else None


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
    for name in ("camera", "intrinsics"):          # UniDepthV2.infer(camera=...), UniDepthV1.infer(intrinsics=...)
        if kw.get(name, None) is not None:
            kw[name] = shard_camera(kw[name], n, lo, hi)
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
