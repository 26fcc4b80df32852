"""Multi-GPU plumbing: frames are independent units, so a batch is split into contiguous blocks,
one per rank (one process per GPU), every rank runs the whole front end on its block with its own handles, and the
fixed-size per-frame result slots are collated with one all_gather per output array (NCCL over NVLink on GPUs; the same
code runs over gloo on CPU tensors in the tests).  No other collective exists on this path.
"""
import torch
import torch.distributed as dist


def shard_range(n_frames, world, rank):
    """contiguous block [start, stop) of rank `rank`; the first n_frames % world ranks get one extra frame"""
    base, extra = divmod(int(n_frames), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n_frames, world):
    return [shard_range(n_frames, world, r)[1] - shard_range(n_frames, world, r)[0] for r in range(world)]


def collate(local, n_frames, group=None):
    """local: dict name -> tensor [n_local, ...] holding this rank's slots (all ranks must pass the same names/dtypes).
    Returns dict name -> tensor [n_frames, ...] in global frame order, on every rank (all_gather semantics)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return dict(local)
    sizes = shard_sizes(n_frames, world)
    pad_to = max(sizes)
    out = {}
    for name, t in local.items():
        n_local = t.shape[0]
        if n_local < pad_to:                       # all_gather needs equal shapes: pad the short shards
            pad = torch.zeros((pad_to - n_local,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            t = torch.cat([t, pad], 0)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous(), group=group)
        out[name] = torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)
    return out


def collate_into(local, out, group=None, async_op=False):
    """Equal-sized shards (the batch divides evenly, as in weak scaling): one all_gather_into_tensor per output array straight into
    the preallocated `out[name]` of shape [world * n_local, ...] -- rank-major order IS global frame order, so there is no temporary
    and no concatenation.  With async_op=True the collectives are only enqueued (they wait for the work already on the current
    stream and then overlap whatever is enqueued next); the returned handles' .wait() makes the current stream wait for them."""
    works = []
    for name, t in local.items():
        w = dist.all_gather_into_tensor(out[name].view(-1), t.contiguous().view(-1), group=group, async_op=async_op)
        if async_op:
            works.append(w)
    return works


def alloc_collated(local, world):
    return {name: torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for name, t in local.items()}
