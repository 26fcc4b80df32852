"""GPU: EVERY frame of a full-size batch against the CPU oracles (the round-1 verdict asked for more than spot checks), the capacity limits that
have no counterpart in the reference (each must surface as B200_ECAPACITY, never as a silent truncation), the match-rich matcher case, the tensor-core
distance stage against the popcount one, two asynchronous matcher calls of one thread on two streams, and the library's own NCCL communicator."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from orb_slam2_aruco_b200 import _lib, synth
from orb_slam2_aruco_b200._lib import B200Error, KP_DTYPE, MARKER_DTYPE, check, lib, ptr
from orb_slam2_aruco_b200.api import FrontEnd, MarkerDetector, ORBextractor, ORBmatcher

pytestmark = pytest.mark.gpu
NCPU = max(1, min(16, os.cpu_count() or 1))


def full_check(imgs, nfeatures, out, rd, rk):
    """all frames: keypoint records as raw bits, descriptor bits, marker ids / corners, match indices"""
    n, h, w = imgs.shape
    k, d, cnt = oracle.orb_extract_batch(imgs, nfeatures, nthreads=NCPU)
    mk, mc = oracle.aruco_detect_batch(imgs, "ARUCO_MIP_25h7", cap=64, nthreads=NCPU)
    assert np.array_equal(out["counts"], cnt)
    assert np.array_equal(out["marker_counts"], mc)
    nmatch = 0
    for f in range(n):
        c = int(cnt[f])
        assert out["kps"][f, :c].tobytes() == np.ascontiguousarray(k[f, :c]).tobytes(), f
        assert np.array_equal(out["desc"][f, :c], d[f, :c]), f
        m = int(mc[f])
        assert np.array_equal(out["markers"][f, :m]["id"], mk[f, :m]["id"]), f
        if m:
            assert np.abs(out["markers"][f, :m]["xy"] - mk[f, :m]["xy"]).max() <= 1e-4, f
        nm, mm = oracle.search_by_bow_bf(rd, rk["angle"], d[f, :c], k[f, :c]["angle"], 0.7, True)
        assert int(out["n_matches"][f]) == nm and np.array_equal(out["matches"][f, :c], mm), f
        nmatch += nm
    return nmatch


def test_every_frame_of_a_c3_batch(built_lib):
    """BASELINE configs[2]: 256 x 640 x 480, 1000 features, 20 markers, match against 1000 descriptors - all 256 frames, no sampling"""
    imgs = synth.make_batch(256, markers=20, first=7000)
    ex = ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, 256)
    fe = FrontEnd(ex, MarkerDetector("ARUCO_MIP_25h7", 640, 480, 256), ORBmatcher(0.7, True))
    rk, rd = ex(synth.make_view(imgs[0], 0, rot_deg=3.0, shift=(5.0, 3.0), noise_sigma=0.0))
    rd, rk = np.ascontiguousarray(rd[:1000]), np.ascontiguousarray(rk[:1000])
    out = fe.process_batch(imgs, rd, rk)
    assert full_check(imgs, 1000, out, rd, rk) > 200


def test_every_frame_of_a_match_rich_batch(built_lib):
    """every frame a perturbed view of the reference scene (rotation, shift, noise): hundreds of matches per frame drive the greedy commit / re-evaluate /
    rescan loop of k_match_resolve and all 30 rotation bins, which the shifted-copy batches never do"""
    scene = synth.make_frame(0, 640, 480, 20)
    imgs = np.stack([synth.make_view(scene, 900 + i) for i in range(48)])
    ex = ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, 48)
    fe = FrontEnd(ex, MarkerDetector("ARUCO_MIP_25h7", 640, 480, 48), ORBmatcher(0.7, True))
    rk, rd = ex(synth.make_view(scene, 0, rot_deg=3.0, shift=(5.0, 3.0), noise_sigma=0.0))
    rd, rk = np.ascontiguousarray(rd[:1000]), np.ascontiguousarray(rk[:1000])
    out = fe.process_batch(imgs, rd, rk)
    assert full_check(imgs, 1000, out, rd, rk) > 48 * 150


def test_every_frame_of_a_c4_shard(built_lib):
    """BASELINE configs[3] geometry: 1280 x 720, 2000 features; 24 distinct frames (the oracle takes ~0.2 s per frame)"""
    imgs = synth.make_batch(24, 1280, 720, markers=20, first=8000)
    ex = ORBextractor(2000, 1.2, 8, 20, 7, 1280, 720, 24)
    fe = FrontEnd(ex, MarkerDetector("ARUCO_MIP_25h7", 1280, 720, 24), ORBmatcher(0.7, True))
    rk, rd = ex(synth.make_view(imgs[0], 0, rot_deg=3.0, shift=(5.0, 3.0), noise_sigma=0.0))
    rd, rk = np.ascontiguousarray(rd[:1000]), np.ascontiguousarray(rk[:1000])
    out = fe.process_batch(imgs, rd, rk)
    full_check(imgs, 2000, out, rd, rk)


def test_streamed_batch_larger_than_the_handles(built_lib):
    """b200_frontend_host with more frames than the handles hold: the batch streams through two alternating scratch regions (C5 runs this way)"""
    base = synth.make_batch(6, markers=20, first=8100)
    imgs = np.concatenate([base] * 50)                       # 300 frames through 256-slot handles: chunks of 32, 96, 128, 44
    ex = ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, 256)
    fe = FrontEnd(ex, MarkerDetector("ARUCO_MIP_25h7", 640, 480, 256), ORBmatcher(0.7, True))
    rk, rd = ex(np.roll(base[1], (3, 5), axis=(0, 1)))
    out = fe.process_batch(imgs, rd, rk)
    assert ex._dims[2] == 256
    full_check(base, 1000, {k: v[:6] for k, v in out.items()}, rd, rk)
    for rep in range(1, 50):
        for key in ("counts", "marker_counts", "n_matches"):
            assert np.array_equal(out[key][6 * rep:6 * rep + 6], out[key][:6]), (key, rep)
        assert np.array_equal(out["desc"][6 * rep:6 * rep + 6], out["desc"][:6]) and np.array_equal(out["matches"][6 * rep:6 * rep + 6], out["matches"][:6])


def test_capacity_limits_fail_loudly(built_lib):
    """limits without a reference counterpart: more than 64 markers per frame, more than 256 quad candidates -> B200_ECAPACITY, not a truncated answer"""
    # 8 x 9 = 72 small markers of ARUCO (ids 0..71)
    img = np.full((480, 640), 200, np.uint8)
    cellpx = 5
    for i in range(72):
        cells = synth.marker_cells("ARUCO", i)
        tile = np.kron(np.where(cells > 0, 235, 20).astype(np.uint8), np.ones((cellpx, cellpx), np.uint8))
        y, x = 14 + (i // 9) * 56, 20 + (i % 9) * 66
        img[y:y + tile.shape[0], x:x + tile.shape[1]] = tile
    assert len(oracle.aruco_detect(img, "ARUCO", cap=256)) == 72
    det = MarkerDetector("ARUCO", 640, 480, 1)
    with pytest.raises(B200Error) as e:
        det.detect(img)
    assert e.value.code == _lib.ECAPACITY and "64 markers" in str(e.value)
    assert len(det.detect(img[:, :330].copy())) > 20                 # the handle keeps working afterwards
    # > 256 convex quads: a fine checkerboard of separate dark squares (each a quad candidate, none a marker)
    img2 = np.full((480, 640), 220, np.uint8)
    for yy in range(6, 470, 26):
        for xx in range(6, 630, 26):
            img2[yy:yy + 20, xx:xx + 20] = 30
    with pytest.raises(B200Error) as e:
        det.detect(img2)
    assert e.value.code == _lib.ECAPACITY
    det.close()
    # a batch larger than a handle that cannot stream it (detector smaller than one pipeline chunk)
    ex = ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, 4)
    with pytest.raises(B200Error) as e:
        check(lib().b200_orb_extract(ex._h, None, 8, 640, 480, 640, 640 * 480, None, None, None, None))
    assert e.value.code == _lib.ECAPACITY


def test_all_dicts_default_is_rejected_with_a_clear_message(built_lib):
    det = MarkerDetector()                                             # the reference's default-constructed detector searches several dictionaries
    with pytest.raises(B200Error) as e:
        det.detect(np.zeros((480, 640), np.uint8))
    assert "setDictionary" in str(e.value)
    det.setDictionary("ARUCO_MIP_36h12")
    assert det.detect(np.zeros((480, 640), np.uint8)) == []


def test_hamming_matrix_with_more_rows_than_grid_y_allows(built_lib):
    """70000 rows: beyond the 65535 limit of grid.y that the round-1 kernel used for the row index"""
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, (70000, 32), dtype=np.uint8); b = rng.integers(0, 256, (3, 32), dtype=np.uint8)
    got = ORBmatcher().distance_matrix(a, b)
    for i in (0, 1, 65535, 65536, 69999):
        for j in range(3):
            assert got[i, j] == int(np.unpackbits(a[i] ^ b[j]).sum())


def test_calls_leave_the_callers_current_device_alone(built_lib):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    torch.cuda.set_device(1)
    ex = ORBextractor(500, 1.2, 8, 20, 7, 640, 480, 1, device=0)
    ex(synth.make_frame(1))
    assert torch.cuda.current_device() == 1
    torch.cuda.set_device(0)


def test_tensor_core_and_popcount_distance_stages_agree(built_lib):
    """k_match_mma (tcgen05, the default) and k_match_topk (B200_MATCH_POPC=1) must leave identical match indices: run the second in a child process"""
    import subprocess
    import sys
    code = ("import numpy as np, sys\n"
            "sys.path.insert(0, %r)\n"
            "from orb_slam2_aruco_b200 import synth\n"
            "from orb_slam2_aruco_b200.api import ORBextractor, ORBmatcher\n"
            "scene = synth.make_frame(0, 640, 480, 20)\n"
            "ex = ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, 1)\n"
            "rk, rd = ex(scene)\n"
            "k, d = ex(synth.make_view(scene, 5))\n"
            "for ratio in (0.6, 0.7, 0.9):\n"
            "    n, m = ORBmatcher(ratio, True).SearchByBoW(rd, rk['angle'], d, k['angle'])\n"
            "    print(n, int(np.dot(m.astype(np.int64) + 2, np.arange(len(m)) + 1)))\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for env in ({}, {"B200_MATCH_POPC": "1"}):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, **env), timeout=300)
        assert r.returncode == 0, r.stderr[-500:]
        outs.append(r.stdout)
    assert outs[0] == outs[1] and len(outs[0].split()) == 6 and int(outs[0].split()[0]) > 100


def test_two_async_matcher_calls_of_one_thread_on_two_streams(built_lib):
    """the device-pointer matcher used ONE thread-local top-K scratch for every stream; two calls in flight on different streams raced on it"""
    import torch
    scene = synth.make_frame(0, 640, 480, 20)
    ex = ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, 1)
    rk, rd = ex(scene)
    views = [ex(synth.make_view(scene, 70 + i)) for i in range(2)]
    want = [oracle.search_by_bow_bf(rd, rk["angle"], d, k["angle"], 0.7, True) for k, d in views]
    dev = torch.device("cuda", 0)
    d_rd = torch.from_numpy(np.ascontiguousarray(rd)).to(dev)
    d_rk = torch.from_numpy(np.ascontiguousarray(rk).view(np.uint8).reshape(-1, 28).copy()).to(dev)
    cap = ex.cap
    bufs = []
    for k, d in views:
        kd = np.zeros((1, cap, 28), np.uint8); dd = np.zeros((1, cap, 32), np.uint8)
        kd[0, :len(k)] = np.ascontiguousarray(k).view(np.uint8).reshape(-1, 28); dd[0, :len(d)] = d
        bufs.append((torch.from_numpy(kd).to(dev), torch.from_numpy(dd).to(dev), torch.tensor([len(k)], dtype=torch.int32, device=dev),
                     torch.zeros((1, cap), dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)))
    streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    m = ORBmatcher(0.7, True)
    for rep in range(20):
        for (kd, dd, nf, mo, nm), s in zip(bufs, streams):
            m.SearchByBoW_device(d_rd, d_rk, len(rd), dd, kd, nf, mo, nm, s)
        torch.cuda.synchronize(dev)
        for (kd, dd, nf, mo, nm), (n2, m2) in zip(bufs, want):
            assert int(nm[0]) == n2 and np.array_equal(mo[0, :len(m2)].cpu().numpy(), m2), rep


def test_library_nccl_communicator_single_rank(built_lib):
    """b200_collate_* with a one-rank communicator: the gather is the root's own device copy; the N > 1 path is exercised by bench.py --gpus N, whose
    rank 0 checks the collated bytes of other ranks against the oracle before printing anything"""
    import torch
    from orb_slam2_aruco_b200 import shard
    if lib().b200_collate_nccl_version() == 0:
        pytest.skip("NCCL not loadable")
    c = shard.Collator(0, 1, 0, shard.Collator.unique_id())
    dev = torch.device("cuda", 0)
    a = torch.arange(1000, dtype=torch.int32, device=dev); b = torch.arange(77, dtype=torch.uint8, device=dev)
    ra = torch.zeros_like(a); rb = torch.zeros_like(b)
    c.gather([a, b], [ra, rb], 0, torch.cuda.current_stream(dev))
    torch.cuda.synchronize(dev)
    assert torch.equal(a, ra) and torch.equal(b, rb) and c.traffic() == (0, 0)
    c.close()
