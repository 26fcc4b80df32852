"""Multi-GPU plumbing: frames are independent units, so a batch is split into contiguous blocks,
one per rank (one process per GPU), every rank runs the whole front end on its block with its own handles, and the
fixed-size per-frame result slots travel to the consumer rank.  No other collective exists on this path.

Two layers:
  * SlotPack + Collator: the product path.  All result arrays of a rank live in ONE packed device buffer (a "bulk" part that is final when the
    extractor has finished: keypoints, descriptors, counts; and a "tail": markers, matches and their counts), and the library's own NCCL
    communicator (csrc/collate.cu, b200_collate_*) moves each part to rank 0 with a grouped ncclSend / ncclRecv: two transfers per step, and
    only the consumer receives anything.
  * collate / collate_into / gather_packed: the same index arithmetic over torch.distributed (gloo on CPU tensors in the tests).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from ._lib import KP_DTYPE, MARKER_DTYPE, check, lib

_TORCH_OF = {np.dtype(np.float32): torch.float32, np.dtype(np.int32): torch.int32, np.dtype(np.uint8): torch.uint8}


class SlotPack:
    """The per-frame result slots of one rank in one contiguous byte buffer.  Sections (each 256-byte aligned):
         bulk: kps [n][cap][7] f32 (28-byte cv::KeyPoint records) | desc [n][cap][32] u8 | counts [n] i32
         tail: markers [n][mcap][9] f32 (36-byte records) | marker_counts [n] i32 | matches [n][cap] i32 | n_matches [n] i32
    `views` are torch tensors into the buffer (what the kernels write), `bulk` / `tail` the two byte ranges the collation moves."""
    FIELDS = (("kps", 28, np.float32, 7), ("desc", 32, np.uint8, 32), ("counts", 4, np.int32, 0),
              ("markers", 36, np.float32, 9), ("marker_counts", 4, np.int32, 0), ("matches", 4, np.int32, 0), ("n_matches", 4, np.int32, 0))

    def __init__(self, n, cap, mcap, device="cpu", detector=True, matcher=True):
        self.n, self.cap, self.mcap = int(n), int(cap), int(mcap)
        per = {"kps": cap, "desc": cap, "counts": 1, "markers": mcap if detector else 0, "marker_counts": 1 if detector else 0,
               "matches": cap if matcher else 0, "n_matches": 1 if matcher else 0}
        self.offsets, o = {}, 0
        for name, rec, _, _ in self.FIELDS:
            nbytes = self.n * per[name] * rec
            self.offsets[name] = (o, nbytes)
            o = (o + nbytes + 255) // 256 * 256
            if name == "counts":
                self.bulk_bytes = o
        self.total_bytes = o
        self.buf = torch.zeros(max(self.total_bytes, 256), dtype=torch.uint8, device=device)
        self.views = self.views_of(self.buf)

    def views_of(self, buf):
        """name -> typed view of a packed buffer of this layout (the rank's own, or one rank's block of the root's collated buffer)"""
        out = {}
        for name, rec, dt, inner in self.FIELDS:
            o, nbytes = self.offsets[name]
            if not nbytes:
                continue
            t = buf[o:o + nbytes].view(_TORCH_OF[np.dtype(dt)])
            per = nbytes // rec // self.n
            out[name] = t.view(self.n, per, inner) if inner else (t.view(self.n, per) if per > 1 else t.view(self.n))
        return out

    @property
    def bulk(self):
        return self.buf[:self.bulk_bytes]

    @property
    def tail(self):
        return self.buf[self.bulk_bytes:self.total_bytes]

    def numpy_views(self, host_bytes):
        """structured numpy views (KP_DTYPE / MARKER_DTYPE records) of a HOST copy of a packed buffer"""
        out = {}
        for name, rec, dt, inner in self.FIELDS:
            o, nbytes = self.offsets[name]
            if not nbytes:
                continue
            raw = host_bytes[o:o + nbytes]
            if name == "kps":
                out[name] = raw.view(KP_DTYPE).reshape(self.n, -1)
            elif name == "markers":
                out[name] = raw.view(MARKER_DTYPE).reshape(self.n, -1)
            elif name == "desc":
                out[name] = raw.reshape(self.n, -1, 32)
            else:
                v = raw.view(dt)
                out[name] = v.reshape(self.n, -1) if len(v) > self.n else v
        return out


class Collator:
    """The library's NCCL communicator (b200_collate_*): gather-to-root of device buffers, all buffers of a call in one NCCL group."""

    def __init__(self, rank, world, device, id_bytes):
        self.rank, self.world = int(rank), int(world)
        h = C.c_void_p()
        idb = (C.c_uint8 * 128).from_buffer_copy(bytes(id_bytes))
        check(lib().b200_collate_create(C.byref(h), idb, self.rank, self.world, int(device)))
        self._h = h

    @staticmethod
    def unique_id():
        idb = (C.c_uint8 * 128)()
        check(lib().b200_collate_unique_id(idb))
        return bytes(idb)

    @classmethod
    def from_torch_distributed(cls, device):
        """rank 0 draws the NCCL id, torch.distributed (whatever its backend) hands it to every rank"""
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(rank, world, device, box[0])

    def gather(self, send, recv=None, root=0, stream=None):
        """send: list of contiguous device tensors (same sizes on every rank); recv (root only): list of tensors of world x the size.  Enqueues only."""
        n = len(send)
        sp = (C.c_void_p * n)(*[t.data_ptr() for t in send])
        rp = (C.c_void_p * n)(*[(t.data_ptr() if t is not None else None) for t in (recv or [None] * n)])
        nb = (C.c_int64 * n)(*[t.numel() * t.element_size() for t in send])
        s = stream.cuda_stream if hasattr(stream, "cuda_stream") else stream
        check(lib().b200_collate_gather(self._h, n, sp, rp if recv is not None else None, nb, int(root), s))

    def traffic(self):
        a, b = C.c_int64(0), C.c_int64(0)
        check(lib().b200_collate_traffic(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def close(self):
        if self._h is not None:
            lib().b200_collate_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gather_packed(part, root_buf, root=0, group=None, async_op=False):
    """torch.distributed form of one SlotPack transfer (gloo on CPU in the tests): `part` = this rank's byte range, root_buf (root only) =
    [world, len(part)] receiving every rank's range; nobody but the root receives anything."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    glist = [root_buf[r] for r in range(world)] if rank == root else None
    return dist.gather(part, gather_list=glist, dst=root, group=group, async_op=async_op)


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
