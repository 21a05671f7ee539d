"""Host-side logic of the image-wise data-parallel path on CPU: shard bounds, pack/unpack, and the
single all-gather over a world_size-2 gloo group."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_out(b, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    return {"confidence": r(b, 1, 6, 8), "intrinsics": r(b, 3, 3), "radius": r(b, 1, 6, 8), "depth": r(b, 1, 6, 8),
            "points": r(b, 3, 6, 8), "rays": r(b, 3, 6, 8), "depth_features": r(b, 2, 3, 16).permute(0, 3, 1, 2)}


def test_shard_bounds_cover_batch():
    from unidepth_b200.parallel import shard_bounds
    for n in (1, 7, 8, 64, 65):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


def test_pack_unpack_round_trip():
    from unidepth_b200.parallel import pack_outputs, unpack_outputs
    out = _fake_out(3, 0)
    back = unpack_outputs(pack_outputs(out), out)
    for k in out:
        assert torch.equal(back[k], out[k]), k


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unidepth_b200.parallel import gather_outputs
    out = _fake_out(2, rank)
    full = gather_outputs(out, world)
    pend = gather_outputs(out, world, async_op=True)      # pipelined form: handle now, dict at wait()
    full2 = pend.wait()
    ok = True
    for r in range(world):
        exp = _fake_out(2, r)
        for k in exp:
            ok &= torch.equal(full[k][2 * r:2 * r + 2], exp[k])
            ok &= torch.equal(full2[k][2 * r:2 * r + 2], exp[k])
    # uneven tail: 5 images over 2 ranks -> 3 + 2
    from unidepth_b200.parallel import shard_bounds
    counts = [shard_bounds(5, r, world)[1] - shard_bounds(5, r, world)[0] for r in range(world)]
    mine = _fake_out(counts[rank], 10 + rank)
    full3 = gather_outputs(mine, world, counts=counts)
    off = 0
    for r in range(world):
        exp = _fake_out(counts[r], 10 + r)
        for k in exp:
            ok &= torch.equal(full3[k][off:off + counts[r]], exp[k])
        off += counts[r]
    ok &= full3["depth"].shape[0] == 5
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gather_outputs_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_camera_and_v1_output_subset():
    """Per-image camera arguments are sliced like the images (a single camera is shared); the gather packs whatever subset of
    the output keys a model returns (UniDepthV1: intrinsics, depth, points)."""
    import pytest
    from unidepth_b200.parallel import pack_outputs, shard_camera, unpack_outputs
    K = torch.arange(6 * 9, dtype=torch.float32).reshape(6, 3, 3)
    assert torch.equal(shard_camera(K, 6, 2, 4), K[2:4])
    assert shard_camera(K[:1], 6, 2, 4) is not None and shard_camera(K[:1], 6, 2, 4).shape[0] == 1
    with pytest.raises(ValueError):
        shard_camera(K[:5], 6, 2, 4)

    class Cams(list):
        pass
    cams = Cams(range(6))
    assert shard_camera(cams, 6, 1, 3) == [1, 2]
    v1 = {"intrinsics": torch.randn(3, 3, 3), "depth": torch.randn(3, 1, 4, 5), "points": torch.randn(3, 3, 4, 5)}
    back = unpack_outputs(pack_outputs(v1), v1)
    assert set(back) == set(v1) and all(torch.equal(back[k], v1[k]) for k in v1)
