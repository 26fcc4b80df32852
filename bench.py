#!/usr/bin/env python
"""bench.py -- frames/s of the B200 front end on BASELINE.json's workloads.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C3|C2]

A "step" is one pass of the hot path over one batch of synthetic frames:
  C3 (default; BASELINE.json configs[2], the configuration the metric "frames/sec (extract+aruco+match)" names):
     batch=256 640x480, ORB extract (1000 features, 8 levels) + ArUco detect (ARUCO_MIP_25h7, 20 planted markers)
     + brute-force SearchByBoW match against a 1000-descriptor reference set
  C2 (configs[1]): the same frames, extract only
`value` is measured with the batch already resident in HBM (CUDA events on the launching stream, one event pair
per step, L2 flushed between steps); `e2e` goes through the reference-facing host-pointer C-ABI call
(b200_frontend_host) with pinned host buffers, H2D and D2H inside the timed region.  Multi-GPU (torchrun, one rank
per GPU): frames are independent, every rank processes its own batch (weak scaling) and the fixed-size result slots
are collated with NCCL all_gather inside the timed region; device time, max over ranks.

--impl reference times the reference's CPU implementation on the same workload with all host threads: the
extractor is the reference's own src/ORBextractor.cc (oracle/_ref, compiled unmodified on the cv shim; the oracle
port when that library is absent), detector and matcher are the oracle restatements (their sources need OpenCV
C++ libraries / Eigen that do not exist in this image).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

DICT = "ARUCO_MIP_25h7"
WORKLOADS = {
    "C2": dict(name="C2: batch=256 synthetic 640x480 gray frames, 1000 ORB features, 8 levels x1.2, FAST 20/7, extract-only",
               metric="frames/sec (extract)", batch=256, w=640, h=480, nfeatures=1000, markers=0, match=False),
    "C3": dict(name="C3: batch=256 synthetic 640x480 gray frames, extract (1000 feat, 8 levels) + ArUco (ARUCO_MIP_25h7, 20 markers/frame) "
                    "+ brute-force match vs 1000-descriptor reference set",
               metric="frames/sec (extract+aruco+match)", batch=256, w=640, h=480, nfeatures=1000, markers=20, match=True),
}
WORKLOADS["C4"] = dict(name="C4: 256 of the 2048 synthetic 1280x720 frames per GPU, extract (2000 feat, 8 levels) + ArUco (ARUCO_MIP_25h7, 20 markers/frame) "
                            "+ brute-force match vs 1000-descriptor reference set",
                       metric="frames/sec (extract+aruco+match)", batch=256, w=1280, h=720, nfeatures=2000, markers=20, match=True, frames_distinct=32)
WORKLOADS["C5"] = dict(name="C5: 128 of the 8192 synthetic 1920x1080 frames per GPU, extract (4000 feat, 8 levels) + ArUco + brute-force match",
                       metric="frames/sec (extract+aruco+match)", batch=128, w=1920, h=1080, nfeatures=4000, markers=20, match=True, frames_distinct=16)
SIGMA_P = {(640, 480): 950532, (1280, 720): 2853088, (1920, 1080): 6419321}      # pixels over the 8 ORB levels


def frames_for(wl, rank, batch=None):
    """deterministic synthetic frames; cached under /tmp because numpy generation takes ~50-150 ms per frame"""
    from orb_slam2_aruco_b200 import synth
    n = batch or wl["batch"]
    path = "/tmp/b200_frames_v2_%dx%d_m%d_r%d_n%d.npy" % (wl["w"], wl["h"], wl["markers"], rank, n)
    if os.path.exists(path):
        try:
            return np.load(path)
        except Exception:
            pass
    distinct = min(n, wl.get("frames_distinct", n))          # the large workloads repeat a few distinct frames (numpy generation is slow)
    imgs = synth.make_batch(distinct, wl["w"], wl["h"], wl["markers"], DICT, first=rank * 100000)
    if distinct < n:
        imgs = np.concatenate([imgs] * ((n + distinct - 1) // distinct))[:n]
    try:
        np.save(path, imgs)
    except Exception:
        pass
    return imgs


def reference_scene(wl):
    """the frame the match reference set is extracted from: frame 0's scene shifted by (5, 3) px"""
    from orb_slam2_aruco_b200 import synth
    return np.roll(synth.make_frame(0, wl["w"], wl["h"], wl["markers"], DICT), (3, 5), axis=(0, 1))


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed region: NVML in-process (a sample every ~2 ms), nvidia-smi as a fallback"""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.stop_flag, self.rows = gpu, False, []
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(gpu))
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(gpu):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v for v in vis.split(",") if v.strip() != ""]
            if gpu < len(ids) and ids[gpu].strip().isdigit():
                return int(ids[gpu])
        return gpu

    def run(self):
        if self.nvml is not None:
            n = self.nvml
            bits = {"hw_slowdown": n.nvmlClocksEventReasonHwSlowdown if hasattr(n, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                    "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
            while not self.stop_flag:
                try:
                    sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                    r = int(get_reasons(self.handle))
                    self.rows.append([str(sm), str(self.sm_max)] + [("Active" if r & bits[k] else "Not Active") for k in
                                                                       ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
                except Exception:
                    pass
                time.sleep(0.002)
            return
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, wl, batch):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture of this workload
    (profiles/ncu_traffic.json, written by tools/ncu_traffic.py); None when no capture matches"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        e = t[kernel]
        if e["workload"] == wl["name"].split(":")[0] and e["batch"] == batch:
            return int(e["dram_bytes_read"] + e["dram_bytes_write"])
    except Exception:
        pass
    return None


def cpu_reference_run(imgs, wl, nthreads, ref_set=None):
    """the reference CPU path on `imgs` with nthreads threads; returns (seconds, kind, description)"""
    import oracle
    n, h, w = imgs.shape
    cap = wl["nfeatures"] + 200
    r = oracle.ref()
    vp = C.c_void_p
    t0 = time.perf_counter()
    if r is not None:
        raw = np.zeros((n, cap, 7), np.float32); desc = np.zeros((n, cap, 32), np.uint8); cnt = np.zeros(n, np.int32)
        r.ref_orb_extract_batch(imgs.ctypes.data_as(vp), n, w, h, w, C.c_long(w * h), wl["nfeatures"], C.c_float(1.2), 8, 20, 7,
                                raw.ctypes.data_as(vp), desc.ctypes.data_as(vp), cnt.ctypes.data_as(vp), cap, nthreads)
        kps28 = np.zeros((n, cap, 28), np.uint8)
        kps28.view(np.float32).reshape(n, cap, 7)[:, :, :5] = raw[:, :, :5]
        kind, what = "reference", "extractor = reference src/ORBextractor.cc on the cv shim (oracle/_ref)"
    else:
        k, desc, cnt = oracle.orb_extract_batch(imgs, wl["nfeatures"], nthreads=nthreads)
        kps28 = np.ascontiguousarray(k).view(np.uint8).reshape(n, -1, 28)
        cap = kps28.shape[1]
        kind, what = "port", "extractor = oracle port"
    if wl["markers"]:
        ra_lib = oracle.ref_aruco()
        if ra_lib is not None:
            mk = np.zeros((n, 256), oracle.MARKER_DTYPE); mc = np.zeros(n, np.int32)
            ra_lib.ref_aruco_detect_batch(imgs.ctypes.data_as(vp), n, w, h, w, C.c_long(w * h), DICT.encode(), mk.ctypes.data_as(vp), mc.ctypes.data_as(vp),
                                          256, nthreads)
            what += "; detector = reference Thirdparty/aruco markerdetector_impl.cpp & co. on the cv shim (oracle/_ref)"
        else:
            oracle.aruco_detect_batch(imgs, DICT, nthreads=nthreads)
            what += "; detector = oracle restatement"
            kind = "port"
    if wl["match"] and ref_set is not None:
        rd, ra = ref_set
        m = np.zeros((n, cap), np.int32); nm = np.zeros(n, np.int32)
        rm = oracle.ref_match()
        if rm is not None and hasattr(rm, "ref_search_by_bow_bf_batch"):
            rm.ref_search_by_bow_bf_batch(rd.ctypes.data_as(vp), ra.ctypes.data_as(vp), len(rd), desc.ctypes.data_as(vp), kps28.ctypes.data_as(vp),
                                          cnt.ctypes.data_as(vp), n, cap, C.c_float(0.7), 1, m.ctypes.data_as(vp), nm.ctypes.data_as(vp), nthreads)
            what += "; matcher = reference src/ORBmatcher.cc SearchByBoW on stand-in headers (oracle/_ref)"
        else:
            oracle.lib().oracle_search_by_bow_bf_batch(rd.ctypes.data_as(vp), ra.ctypes.data_as(vp), len(rd), desc.ctypes.data_as(vp),
                                                       kps28.ctypes.data_as(vp), cnt.ctypes.data_as(vp), n, cap, C.c_float(0.7), 1,
                                                       C.c_float(np.float32(30.0) / np.float32(360.0)), m.ctypes.data_as(vp), nm.ctypes.data_as(vp), nthreads)
            what += "; matcher = oracle restatement"
        kind = "port" if kind == "port" else "reference"
    return time.perf_counter() - t0, kind, what


def cpu_ref_set(wl):
    import oracle
    k, d = oracle.orb_extract(reference_scene(wl), wl["nfeatures"])
    return np.ascontiguousarray(d[:1000]), np.ascontiguousarray(k["angle"][:1000])


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    try:
        ncores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    sample = wl["batch"]                                   # one step = the whole batch (about a second on 16 host threads)
    imgs = frames_for(wl, 0)[:sample]
    ref_set = cpu_ref_set(wl) if wl["match"] else None
    for _ in range(args.warmup):
        cpu_reference_run(imgs, wl, ncores, ref_set)
    t, kind, what = 0.0, "port", ""
    for _ in range(args.steps):
        dt, kind, what = cpu_reference_run(imgs, wl, ncores, ref_set)
        t += dt
    fps = sample * args.steps / t
    line = {"impl": "reference", "metric": wl["metric"], "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": {"workload": wl["name"]},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": ncores, "kind": kind,
                             "sample": "%d of the %d frames per step, %d host threads; %s" % (sample, wl["batch"], ncores, what)},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-API leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    import torch.distributed as dist
    from orb_slam2_aruco_b200 import _lib
    from orb_slam2_aruco_b200.api import FrontEnd, MarkerDetector, ORBextractor, ORBmatcher

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B, W, H = wl["batch"], wl["w"], wl["h"]
    imgs_np = frames_for(wl, rank)
    ex = ORBextractor(wl["nfeatures"], 1.2, 8, 20, 7, W, H, B, device=local)
    det = MarkerDetector(DICT, W, H, B, device=local) if wl["markers"] else None
    matcher = ORBmatcher(0.7, True, device=local) if wl["match"] else None
    cap, mcap = ex.cap, 64
    d_imgs = torch.from_numpy(imgs_np).to(dev)
    d_kps = torch.zeros((B, cap, 7), dtype=torch.float32, device=dev)
    d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev)
    d_counts = torch.zeros((B,), dtype=torch.int32, device=dev)
    d_markers = torch.zeros((B, mcap, 9), dtype=torch.float32, device=dev)
    d_mcounts = torch.zeros((B,), dtype=torch.int32, device=dev)
    d_match = torch.zeros((B, cap), dtype=torch.int32, device=dev)
    d_nmatch = torch.zeros((B,), dtype=torch.int32, device=dev)
    # reference set for the matcher: <= 1000 descriptors of the shifted scene, extracted with the CUDA extractor
    ref_np = None
    if wl["match"]:
        rk, rd = ex(reference_scene(wl))
        ref_np = (np.ascontiguousarray(rd[:1000]), np.ascontiguousarray(rk[:1000]))
        d_rdesc = torch.from_numpy(ref_np[0]).to(dev)
        d_rkps = torch.from_numpy(ref_np[1].view(np.uint8).reshape(-1, 28).copy()).to(dev)
        n_ref = len(ref_np[0])
    prio = [int(v) for v in os.environ.get("B200_BENCH_PRIO", "0,-1").split(",")]      # detector stream at high priority (latency-bound kernels start early)
    s_main = torch.cuda.Stream(device=dev, priority=prio[0])
    s_aux = torch.cuda.Stream(device=dev, priority=prio[1])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    outs = {"kps": d_kps, "desc": d_desc, "counts": d_counts}
    if det:
        outs.update(markers=d_markers, marker_counts=d_mcounts)
    if matcher:
        outs.update(matches=d_match, n_matches=d_nmatch)
    from orb_slam2_aruco_b200 import shard
    ev_fork, ev_join = torch.cuda.Event(), torch.cuda.Event()

    collated = shard.alloc_collated(outs, world) if world > 1 else {}

    def step():
        works = []
        if det is not None:                      # detector on its own stream, concurrently with extractor + matcher
            ev_fork.record(s_main)
            s_aux.wait_event(ev_fork)
            det.detect_batch_device(d_imgs, d_markers, d_mcounts, s_aux)
            ev_join.record(s_aux)
        ex.extract_batch_device(d_imgs, d_kps, d_desc, d_counts, s_main)
        if world > 1:                            # every rank ends up with all B*world result slots (rank 0 is the consumer): the bulk
            with torch.cuda.stream(s_main):      # (keypoints + descriptors) is gathered while the matcher runs
                works += shard.collate_into({k: outs[k] for k in ("kps", "desc", "counts")}, collated, async_op=True)
        if matcher is not None:
            matcher.SearchByBoW_device(d_rdesc, d_rkps, n_ref, d_desc, d_kps, d_counts, d_match, d_nmatch, s_main)
        if det is not None:
            s_main.wait_event(ev_join)
        if world > 1:
            with torch.cuda.stream(s_main):
                works += shard.collate_into({k: v for k, v in outs.items() if k not in ("kps", "desc", "counts")}, collated, async_op=True)
                for w in works:
                    w.wait()

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    sync_all()
    ex.set_profile(True)
    launches0 = _lib.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stage = np.zeros(4, np.float64)
    sync_all()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        with torch.cuda.stream(s_main):
            flush.fill_(i & 0xff)          # evict the batch from L2 (not timed)
            ev[i][0].record(s_main)
        step()
        ev[i][1].record(s_main)
        s_main.synchronize()
        stage += ex.stage_ms()
        stage_frames = ex.stage_frames()
    sync_all()
    t_wall = time.perf_counter() - t_wall0
    launches = _lib.launch_count() - launches0
    total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    ex.set_profile(False)
    nkp = int(d_counts.sum().item())
    nmk = int(d_mcounts.sum().item()) if det else 0
    nmatch = int(d_nmatch.sum().item()) if matcher else 0

    # ---- end to end through the reference-facing host-pointer C-ABI (pinned buffers) ----------------
    e2e_s = float("nan")
    if not args.no_e2e:
        fe = FrontEnd(ex, det, matcher)
        h_imgs = torch.from_numpy(imgs_np).pin_memory()
        out = fe.alloc_outputs(B, pinned=True)
        for _ in range(2):
            fe.process_batch(h_imgs.numpy(), ref_np[0] if ref_np else None, ref_np[1] if ref_np else None, out=out)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fe.process_batch(h_imgs.numpy(), ref_np[0] if ref_np else None, ref_np[1] if ref_np else None, out=out)
        torch.cuda.synchronize(dev)
        e2e_s = time.perf_counter() - t0
        assert int(out["counts"].sum()) == nkp, "host path and device path disagree"
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    # ---- max over ranks ----------------------------------------------------------------------------
    t = torch.tensor([total_ms, e2e_s * 1000.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])
    frames = B * world * args.steps
    fps = frames / (total_ms / 1000.0)
    e2e_fps = frames / (e2e_ms / 1000.0)

    if rank == 0:
        peak, peak_src = peaks()
        sp = SIGMA_P.get((W, H), int(3.0942 * W * H))
        fast_bytes = (sp + 4 * int(10000 * W * H / 307200)) * stage_frames              # k_fast: every level pixel read once + ~10k candidate slots written / frame,
                                                                  # for the frames of the timed launch (large batches run as two half-batch launches)
        fast_ms = stage[1] / args.steps
        achieved = fast_bytes / (fast_ms / 1000.0) / 1e9
        b_frame = 2 * sp + 60 * (nkp / B)                         # B_ext
        if det:
            b_frame += 3.333 * W * H + 36 * (nmk / B)             # B_aru
        if matcher:
            b_frame += 36 * (nkp / B + 1000) + 4 * (nkp / B)      # B_mat
        d2h = B * cap * 60 + B * 4 + (B * mcap * 36 + B * 4 if det else 0) + (B * cap * 4 + B * 4 if matcher else 0)
        line = {
            "metric": wl["metric"], "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": wl["name"], "frames_per_gpu": B, "l2": "flushed between steps (256 MiB fill, untimed)",
                       "timing": "CUDA events on the launching stream, one pair per step, max over ranks",
                       "streams": "extractor + matcher on the launching stream, detector on a second, higher-priority stream",
                       "collate": "nccl all_gather_into_tensor of the fixed result slots inside the step; keypoints + descriptors gathered while the matcher runs" if world > 1 else "none (1 GPU)"},
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": int(B * W * H), "d2h_bytes_per_step": int(d2h),
                    "api": "b200_frontend_host (pinned host buffers, chunked H2D overlapped with compute)"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "k_fast", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic("k_fast", wl, stage_frames), "peak_source": peak_src, "ms_per_launch": fast_ms,
                         "algorithmic_bytes_per_launch": int(fast_bytes), "frames_per_launch": int(stage_frames),
                         "extractor_stage_ms": {"pyramid": stage[0] / args.steps, "fast": stage[1] / args.steps,
                                                "quadtree": stage[2] / args.steps, "describe": stage[3] / args.steps},
                         "whole_step": {"algorithmic_bytes_per_frame": b_frame, "achieved_gbs": fps / world * b_frame / 1e9,
                                        "frac": fps / world * b_frame / 1e9 / peak}},
            "clocks": sampler.summary(),
            "wall_s_timed_region": t_wall,
            "per_frame": {"keypoints": nkp / B, "markers": nmk / B, "matches": nmatch / B},
        }
        if not args.no_cpu_baseline and world == 1:
            sample = max(16, min(B, 256 * 640 * 480 // (W * H)))    # ~10 s of single-thread CPU work whatever the frame size
            ref_set = cpu_ref_set(wl) if wl["match"] else None
            secs, kind, what = cpu_reference_run(imgs_np[:sample], wl, 1, ref_set)
            line["cpu_baseline"] = {"value": sample / secs, "unit": "frames/s", "cores": 1, "kind": kind,
                                    "sample": "first %d of the %d frames, 1 thread; %s" % (sample, B, what)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
