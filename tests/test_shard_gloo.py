"""CPU, world_size 2 over gloo: the frame sharding + result collation used for N > 1 GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from orb_slam2_aruco_b200 import shard


def test_shard_ranges_cover_the_batch():
    for n in (0, 1, 5, 256, 2048, 8191):
        for world in (1, 2, 4, 8):
            spans = [shard.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == shard.shard_sizes(n, world)


def fake_slots(lo, hi, cap=6):
    """deterministic per-frame result slots as a function of the global frame index"""
    idx = torch.arange(lo, hi)
    counts = (idx % cap + 1).to(torch.int32)
    kps = (idx[:, None, None] * 1000 + torch.arange(cap)[None, :, None] * 10 + torch.arange(7)[None, None, :]).to(torch.float32)
    desc = ((idx[:, None, None] + torch.arange(cap)[None, :, None] * 3 + torch.arange(32)[None, None, :]) % 256).to(torch.uint8)
    return {"counts": counts, "kps": kps, "desc": desc}


def _worker(rank, world, port, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_range(n_frames, world, rank)
    got = shard.collate(fake_slots(lo, hi), n_frames)
    want = fake_slots(0, n_frames)
    ok = all(torch.equal(got[k], want[k]) for k in want)
    if n_frames % world == 0:                    # equal shards: the allocation-free path bench.py uses, synchronous and asynchronous
        local = fake_slots(lo, hi)
        for async_op in (False, True):
            out = shard.alloc_collated(local, world)
            for w in shard.collate_into(local, out, async_op=async_op):
                w.wait()
            ok = ok and all(torch.equal(out[k], want[k]) for k in want)
    q.put((rank, ok, int(got["counts"].shape[0])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [8, 5])
def test_collate_world2_gloo(n_frames):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res) and all(r[2] == n_frames for r in res)


def test_collate_single_process_is_identity():
    x = fake_slots(0, 4)
    y = shard.collate(x, 4)
    assert all(torch.equal(x[k], y[k]) for k in x)


def test_slot_pack_layout():
    """the packed result slots: aligned, non-overlapping sections whose views have the shapes the kernels write"""
    p = shard.SlotPack(4, 1024, 64)
    spans = sorted(p.offsets.values())
    assert all(o % 256 == 0 for o, _ in spans) and all(a[0] + a[1] <= b[0] for a, b in zip(spans, spans[1:]))
    assert p.bulk_bytes == p.offsets["markers"][0] and p.total_bytes >= spans[-1][0] + spans[-1][1]
    assert tuple(p.views["kps"].shape) == (4, 1024, 7) and tuple(p.views["desc"].shape) == (4, 1024, 32) and tuple(p.views["matches"].shape) == (4, 1024)
    p.views["counts"][:] = torch.tensor([1, 2, 3, 4], dtype=torch.int32)
    p.views["markers"][2, 5, 0] = 7.0
    host = p.buf.numpy()
    nv = p.numpy_views(host)
    assert nv["counts"].tolist() == [1, 2, 3, 4] and nv["markers"][2, 5]["id"].view(np.float32) == 7.0 and nv["kps"].shape == (4, 1024)
    q = shard.SlotPack(2, 100, 64, detector=False, matcher=False)            # extract-only: the tail is empty
    assert q.total_bytes == q.bulk_bytes and "markers" not in q.views


def _pack_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, cap = 3, 6
    pack = shard.SlotPack(n, cap, 2)
    slots = fake_slots(rank * n, (rank + 1) * n, cap)
    pack.views["counts"].copy_(slots["counts"]); pack.views["kps"].copy_(slots["kps"]); pack.views["desc"].copy_(slots["desc"])
    pack.views["matches"].fill_(rank + 10)
    root_bulk = torch.zeros((world, pack.bulk_bytes), dtype=torch.uint8) if rank == 0 else None
    root_tail = torch.zeros((world, pack.total_bytes - pack.bulk_bytes), dtype=torch.uint8) if rank == 0 else None
    w1 = shard.gather_packed(pack.bulk, root_bulk, 0, async_op=True)            # two transfers per step, like bench.py
    w2 = shard.gather_packed(pack.tail, root_tail, 0, async_op=True)
    w1.wait(); w2.wait()
    ok = True
    if rank == 0:
        for r in range(world):
            v = pack.views_of(torch.cat([root_bulk[r], root_tail[r]]))
            want = fake_slots(r * n, (r + 1) * n, cap)
            ok = ok and all(torch.equal(v[k], want[k]) for k in want) and bool((v["matches"] == r + 10).all())
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_packed_gather_to_root_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pack_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] for r in res)
