#!/usr/bin/env python
"""bench.py -- frames/s of the B200 front end on BASELINE.json's workloads.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C3|C2]

The ONE JSON line carries the C3 headline (below) and, under `extra_workloads`, the other configurations of BASELINE.json measured in the same
process: C1 (single-frame latency through the C++ adapters; N=1 only), C4 and C5 (per-GPU shards of configs[3] / [4], at every N), and a
match-rich C3 variant (every frame a perturbed view of the reference scene).  At N>1 rank 0 checks the COLLATED buffers of other ranks' frames
against the CPU oracle after the timed loop and refuses to print a line on a mismatch.

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
WORKLOADS["C4"] = dict(name="C4: 256 of the 2048 synthetic 1280x720 frames per GPU (the 8-GPU shard of BASELINE configs[3]), extract (2000 feat, 8 levels) + ArUco "
                            "(ARUCO_MIP_25h7, 20 markers/frame) + brute-force match vs 1000-descriptor reference set",
                       metric="frames/sec (extract+aruco+match)", batch=256, w=1280, h=720, nfeatures=2000, markers=20, match=True, frames_distinct=32)
WORKLOADS["C5"] = dict(name="C5: 1024 of the 8192 synthetic 1920x1080 frames per GPU (the 8-GPU shard of BASELINE configs[4]) in 8 sub-batches of 128, extract "
                            "(4000 feat, 8 levels) + ArUco + brute-force match",
                       metric="frames/sec (extract+aruco+match)", batch=128, sub=8, w=1920, h=1080, nfeatures=4000, markers=20, match=True, frames_distinct=16)
WORKLOADS["C3R"] = dict(name="C3 match-rich: batch=256 640x480, every frame a perturbed view (rotation <= 3 deg, shift <= 8 px, fresh noise) of the scene the "
                             "1000-descriptor reference set comes from (~300 matches per frame instead of ~10), extract + ArUco + brute-force match",
                        metric="frames/sec (extract+aruco+match)", batch=256, w=640, h=480, nfeatures=1000, markers=20, match=True, views=True)
SIGMA_P = {(640, 480): 950532, (1280, 720): 2853088, (1920, 1080): 6419321}      # pixels over the 8 ORB levels


def frames_for(wl, rank, batch=None):
    """deterministic synthetic frames; cached under /tmp because numpy generation takes ~50-150 ms per frame"""
    from orb_slam2_aruco_b200 import synth
    n = batch or wl["batch"]
    path = "/tmp/b200_frames_v3_%dx%d_m%d_r%d_n%d%s.npy" % (wl["w"], wl["h"], wl["markers"], rank, n, "_views" if wl.get("views") else "")
    if os.path.exists(path):
        try:
            return np.load(path)
        except Exception:
            pass
    distinct = min(n, wl.get("frames_distinct", n))          # the large workloads repeat a few distinct frames (numpy generation is slow)
    if wl.get("views"):
        scene = synth.make_frame(0, wl["w"], wl["h"], wl["markers"], DICT)
        imgs = np.stack([synth.make_view(scene, rank * 100000 + i) for i in range(distinct)])
    else:
        imgs = synth.make_batch(distinct, wl["w"], wl["h"], wl["markers"], DICT, first=rank * 100000)
    if distinct < n:
        imgs = np.concatenate([imgs] * ((n + distinct - 1) // distinct))[:n]
    try:
        np.save(path, imgs)
    except Exception:
        pass
    return imgs


def reference_scene(wl):
    """the frame the match reference set is extracted from: frame 0's scene shifted by (5, 3) px and rotated by 3 degrees"""
    from orb_slam2_aruco_b200 import synth
    return synth.make_view(synth.make_frame(0, wl["w"], wl["h"], wl["markers"], DICT), 0, rot_deg=3.0, shift=(5.0, 3.0), noise_sigma=0.0)


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


def cpu_reference_run(imgs, wl, nthreads, ref_set=None, DICT=DICT):
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


class Ctx:
    pass


class Workload:
    """handles, resident inputs, packed result slots and streams of one workload on this rank"""

    def __init__(self, ctx, wl):
        import torch
        from orb_slam2_aruco_b200 import shard
        from orb_slam2_aruco_b200.api import MarkerDetector, ORBextractor, ORBmatcher
        self.ctx, self.wl = ctx, wl
        dev, local, rank, world = ctx.dev, ctx.local, ctx.rank, ctx.world
        self.B, self.W, self.H, self.nsub = wl["batch"], wl["w"], wl["h"], wl.get("sub", 1)
        B = self.B
        self.imgs_np = frames_for(wl, rank)                            # one sub-batch worth of distinct host frames
        self.ex = ORBextractor(wl["nfeatures"], 1.2, 8, 20, 7, self.W, self.H, B, device=local)
        self.det = MarkerDetector(DICT, self.W, self.H, B, device=local) if wl["markers"] else None
        self.matcher = ORBmatcher(0.7, True, device=local) if wl["match"] else None
        self.cap, self.mcap = self.ex.cap, (self.det.cap if self.det else 0)
        one = torch.from_numpy(self.imgs_np).to(dev)
        # all frames of a step resident in HBM: nsub sub-batches (C5: 8 x 128 frames = 2.1 GB; the 16 distinct frames repeat)
        self.d_imgs = one if self.nsub == 1 else one.unsqueeze(0).repeat(self.nsub, 1, 1, 1).reshape(self.nsub * B, self.H, self.W).contiguous()
        self.pack = shard.SlotPack(B, self.cap, max(self.mcap, 1), device=dev, detector=self.det is not None, matcher=self.matcher is not None)
        self.v = self.pack.views
        # sub-batched workloads (C5) alternate between TWO sets of handles, result slots and streams, like b200_frontend_host does with its chunks:
        # the latency-bound tail of one sub-batch (quadtree at 4000 features, contour walks, greedy resolve) runs under the dense kernels of the next
        self.sets = [dict(ex=self.ex, det=self.det, pack=self.pack)]
        if self.nsub > 1 and os.environ.get("B200_BENCH_ONE_SET") is None:
            ex2 = ORBextractor(wl["nfeatures"], 1.2, 8, 20, 7, self.W, self.H, B, device=local)
            det2 = MarkerDetector(DICT, self.W, self.H, B, device=local) if wl["markers"] else None
            pack2 = shard.SlotPack(B, self.cap, max(self.mcap, 1), device=dev, detector=self.det is not None, matcher=self.matcher is not None)
            self.sets.append(dict(ex=ex2, det=det2, pack=pack2))
        self.ref_np = None
        if wl["match"]:          # reference set for the matcher: <= 1000 descriptors of the reference scene, extracted with the CUDA extractor
            rk, rd = self.ex(reference_scene(wl))
            self.ref_np = (np.ascontiguousarray(rd[:1000]), np.ascontiguousarray(rk[:1000]))
            self.d_rdesc = torch.from_numpy(self.ref_np[0]).to(dev)
            self.d_rkps = torch.from_numpy(self.ref_np[1].view(np.uint8).reshape(-1, 28).copy()).to(dev)
            self.n_ref = len(self.ref_np[0])
        prio = [int(v) for v in os.environ.get("B200_BENCH_PRIO", "0,-1").split(",")]      # detector stream at high priority (latency-bound kernels start early)
        for st in self.sets:
            st["s_main"] = torch.cuda.Stream(device=dev, priority=prio[0])
            st["s_aux"] = torch.cuda.Stream(device=dev, priority=prio[1])
            st["ev_fork"], st["ev_join"], st["ev_bulk"], st["ev_tail"], st["ev_done"], st["ev_end"] = (torch.cuda.Event() for _ in range(6))
        self.s_main, self.s_aux = self.sets[0]["s_main"], self.sets[0]["s_aux"]
        self.s_comm = torch.cuda.Stream(device=dev, priority=-1)
        self.ev_begin = torch.cuda.Event()
        self.coll_bulk = self.coll_tail = None
        if world > 1 and rank == 0:      # the consumer's receive buffers: [sub-batch][rank][bytes]
            self.coll_bulk = torch.zeros((self.nsub, world, self.pack.bulk_bytes), dtype=torch.uint8, device=dev)
            self.coll_tail = torch.zeros((self.nsub, world, self.pack.total_bytes - self.pack.bulk_bytes), dtype=torch.uint8, device=dev)

    def step(self):
        """one pass over all sub-batches; enqueued on the sets' streams, begins and ends on self.s_main (the stream the timing events are recorded on)"""
        ctx = self.ctx
        if len(self.sets) > 1:
            self.ev_begin.record(self.s_main)
            self.sets[1]["s_main"].wait_event(self.ev_begin)
        for sub in range(self.nsub):
            st = self.sets[sub % len(self.sets)]
            v, ex, det, s_main, s_aux = st["pack"].views, st["ex"], st["det"], st["s_main"], st["s_aux"]
            imgs = self.d_imgs[sub * self.B:(sub + 1) * self.B]
            if det is not None:                           # detector on its own stream, concurrently with extractor + matcher
                st["ev_fork"].record(s_main)
                s_aux.wait_event(st["ev_fork"])
                det.detect_batch_device(imgs, v["markers"], v["marker_counts"], s_aux)
                st["ev_join"].record(s_aux)
            ex.extract_batch_device(imgs, v["kps"], v["desc"], v["counts"], s_main)
            if ctx.world > 1:                             # bulk (keypoints + descriptors + counts) leaves for rank 0 while the matcher runs
                st["ev_bulk"].record(s_main)
                self.s_comm.wait_event(st["ev_bulk"])
                ctx.collator.gather([st["pack"].bulk], [self.coll_bulk[sub]] if ctx.rank == 0 else None, 0, self.s_comm)
            if self.matcher is not None:
                self.matcher.SearchByBoW_device(self.d_rdesc, self.d_rkps, self.n_ref, v["desc"], v["kps"], v["counts"], v["matches"], v["n_matches"], s_main)
            if det is not None:
                s_main.wait_event(st["ev_join"])
            if ctx.world > 1:                             # tail (markers + matches) after the join; the sub-batch ends when both transfers have landed
                st["ev_tail"].record(s_main)
                self.s_comm.wait_event(st["ev_tail"])
                ctx.collator.gather([st["pack"].tail], [self.coll_tail[sub]] if ctx.rank == 0 else None, 0, self.s_comm)
                st["ev_done"].record(self.s_comm)
                s_main.wait_event(st["ev_done"])
        if len(self.sets) > 1:
            self.sets[1]["ev_end"].record(self.sets[1]["s_main"])
            self.s_main.wait_event(self.sets[1]["ev_end"])

    def totals(self):
        v = self.v
        nkp = int(v["counts"].sum().item())
        nmk = int(v["marker_counts"].sum().item()) if self.det else 0
        nmatch = int(v["n_matches"].sum().item()) if self.matcher else 0
        return nkp, nmk, nmatch


def sync_all(ctx):
    import torch
    torch.cuda.synchronize(ctx.dev)
    if ctx.world > 1:
        ctx.dist.barrier()
        torch.cuda.synchronize(ctx.dev)


def max_over_ranks(ctx, values):
    import torch
    t = torch.tensor(values, dtype=torch.float64, device=ctx.dev)
    if ctx.world > 1:
        ctx.dist.all_reduce(t, op=ctx.dist.ReduceOp.MAX)
    return [float(x) for x in t]


def timed_resident(ctx, wk, steps, warmup, profile=False):
    """W untimed warm-up steps, then exactly `steps` steps, each bracketed by CUDA events on the launching stream with an untimed L2 flush in
    front; barrier + synchronize on both sides; returns (sum of event ms on this rank, launches, extractor stage ms, frames of the stage launch)"""
    import torch
    from orb_slam2_aruco_b200 import _lib
    for _ in range(warmup):
        wk.step()
    sync_all(ctx)
    if profile:
        wk.ex.set_profile(True)
    launches0 = _lib.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    stage, stage_frames = np.zeros(4, np.float64), 0
    sync_all(ctx)
    t0 = time.perf_counter()
    for i in range(steps):
        with torch.cuda.stream(wk.s_main):
            ctx.flush.fill_(i & 0xff)          # evict the batch from L2 (not timed)
            ev[i][0].record(wk.s_main)
        wk.step()
        ev[i][1].record(wk.s_main)
        wk.s_main.synchronize()
        if profile:
            stage += wk.ex.stage_ms()
            stage_frames = wk.ex.stage_frames()
    sync_all(ctx)
    wall = time.perf_counter() - t0
    launches = _lib.launch_count() - launches0
    if profile:
        wk.ex.set_profile(False)
    return float(sum(a.elapsed_time(b) for a, b in ev)), int(launches), stage, stage_frames, wall


def timed_e2e(ctx, wk, steps):
    """the reference-facing host API: pinned host frames in, pinned host results out (b200_frontend_host), and at N > 1 the collation of every
    rank's results into rank 0's HOST buffers (b200_frontend_collate_host) - all inside the timed region.  Returns (seconds, h2d, d2h bytes/step)."""
    import torch
    from orb_slam2_aruco_b200.api import FrontEnd
    fe = FrontEnd(wk.ex, wk.det, wk.matcher)
    B = wk.B
    rd, rk = wk.ref_np if wk.ref_np else (None, None)
    if wk.nsub > 1 and ctx.world == 1:
        # all sub-batches of the step in ONE call: b200_frontend_host pipelines 128-frame chunks (upload / compute / download overlap across
        # the whole step) through two alternating scratch regions, so the handles stay at 256 frame slots
        n_all = B * wk.nsub
        h_all = torch.from_numpy(np.concatenate([wk.imgs_np] * wk.nsub)).pin_memory()
        out = fe.alloc_outputs(n_all, pinned=True)

        def one():
            fe.process_batch(h_all.numpy(), rd, rk, out=out)
        for _ in range(2):
            one()
        sync_all(ctx)
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        torch.cuda.synchronize(ctx.dev)
        secs = time.perf_counter() - t0
        assert int(out["counts"][-B:].sum()) == wk.totals()[0], "host path and device path disagree"
        cap, mcap = wk.cap, wk.mcap
        d2h = n_all * cap * 60 + n_all * 4 + (n_all * mcap * 36 + n_all * 4 if wk.det else 0) + (n_all * cap * 4 + n_all * 4 if wk.matcher else 0)
        return secs, int(n_all * wk.W * wk.H), int(d2h)
    h_imgs = torch.from_numpy(wk.imgs_np).pin_memory()
    out = fe.alloc_outputs(B, pinned=True)
    root_out = fe.alloc_outputs(B * ctx.world, pinned=True) if ctx.world > 1 and ctx.rank == 0 else None

    def one():
        for _ in range(wk.nsub):
            fe.process_batch(h_imgs.numpy(), rd, rk, out=out)
            if ctx.world > 1:
                fe.collate_host(ctx.collator, root_out, 0)
    for _ in range(2):
        one()
    sync_all(ctx)
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize(ctx.dev)
    secs = time.perf_counter() - t0
    nkp = wk.totals()[0]
    assert int(out["counts"].sum()) == nkp, "host path and device path disagree"
    if root_out is not None:
        assert int(root_out["counts"][:B].sum()) == nkp, "collated host buffers do not start with rank 0's results"
    cap, mcap = wk.cap, wk.mcap
    d2h = B * cap * 60 + B * 4 + (B * mcap * 36 + B * 4 if wk.det else 0) + (B * cap * 4 + B * 4 if wk.matcher else 0)
    return secs, int(B * wk.W * wk.H) * wk.nsub, int(d2h) * wk.nsub * (ctx.world if ctx.rank == 0 and ctx.world > 1 else 1)


def check_collated(ctx, wk):
    """N > 1, after the timed loop: what rank 0 RECEIVED is compared (1) byte-for-byte checksums of every rank's packed slots, (2) with the CPU
    oracle on sample frames of OTHER ranks (keypoints, descriptor bits, marker ids / corners, match indices).  Raises on any mismatch: no
    bench line is printed for a collation that delivers wrong data."""
    import torch
    dist = ctx.dist
    sub = wk.nsub - 1                                            # the pack of its set holds the last sub-batch of the last step
    last = wk.sets[sub % len(wk.sets)]["pack"]
    mine = torch.stack([last.bulk.to(torch.int64).sum(), last.tail.to(torch.int64).sum()])
    sums = [torch.zeros_like(mine) for _ in range(ctx.world)]
    dist.all_gather(sums, mine)
    if ctx.rank != 0:
        return None
    import oracle
    for r in range(ctx.world):
        got = torch.stack([wk.coll_bulk[sub, r].to(torch.int64).sum(), wk.coll_tail[sub, r].to(torch.int64).sum()])
        if not torch.equal(got, sums[r]):
            raise SystemExit("collated slots of rank %d differ from what that rank produced (checksum)" % r)
    checked = []
    for r in sorted({1, ctx.world - 1}):
        host = torch.cat([wk.coll_bulk[sub, r], wk.coll_tail[sub, r]]).cpu().numpy()
        o = wk.pack.numpy_views(host)
        frames = frames_for(wk.wl, r)
        for f in sorted({0, wk.B // 2 + 1, wk.B - 1}):
            img = frames[f]
            k2, d2 = oracle.orb_extract(img, wk.wl["nfeatures"])
            n = int(o["counts"][f])
            ok = n == len(k2) and np.array_equal(o["desc"][f, :n], d2) and all(
                np.array_equal(o["kps"][f, :n][name].view(np.uint32), k2[name].view(np.uint32)) for name in k2.dtype.names)
            if ok and wk.det is not None:
                want = oracle.aruco_detect(img, DICT)
                m = int(o["marker_counts"][f])
                ok = m == len(want) and np.array_equal(o["markers"][f, :m]["id"], want["id"]) and (m == 0 or float(np.abs(o["markers"][f, :m]["xy"] - want["xy"]).max()) <= 1e-4)
            if ok and wk.matcher is not None:
                n2, m2 = oracle.search_by_bow_bf(wk.ref_np[0], wk.ref_np[1]["angle"], d2, k2["angle"], 0.7, True)
                ok = int(o["n_matches"][f]) == n2 and np.array_equal(o["matches"][f, :n], m2)
            if not ok:
                raise SystemExit("collated results of rank %d frame %d differ from the CPU oracle" % (r, f))
            checked.append([r, f])
    return {"checksums": "all %d ranks equal" % ctx.world, "oracle_frames_rank_frame": checked, "result": "bit-exact (corners <= 1e-4 px)"}


def workload_line(ctx, wl, steps, warmup, do_e2e=True, headline=False):
    """measure one workload; returns the dict rank 0 prints (None on other ranks)"""
    wk = Workload(ctx, wl)
    B, W, H, nsub = wk.B, wk.W, wk.H, wk.nsub
    traffic_before = ctx.collator.traffic() if ctx.world > 1 else (0, 0)
    total_ms, launches, stage, stage_frames, wall = timed_resident(ctx, wk, steps, warmup, profile=True)
    nkp, nmk, nmatch = wk.totals()
    traffic0 = ctx.collator.traffic() if ctx.world > 1 else (0, 0)
    parity = check_collated(ctx, wk) if ctx.world > 1 else None
    e2e_s, h2d, d2h = float("nan"), 0, 0
    if do_e2e:
        e2e_s, h2d, d2h = timed_e2e(ctx, wk, steps)
    total_ms, e2e_ms = max_over_ranks(ctx, [total_ms, e2e_s * 1000.0])
    if ctx.rank != 0:
        return None
    frames = B * nsub * ctx.world * steps
    fps = frames / (total_ms / 1000.0)
    peak, peak_src = peaks()
    sp = SIGMA_P.get((W, H), int(3.0942 * W * H))
    fast_bytes = (sp + 4 * int(10000 * W * H / 307200)) * stage_frames      # k_fast: every level pixel read once + ~10k candidate slots written / frame,
    fast_ms = stage[1] / steps                                              # for the frames of the timed launch (large batches run as two half-batch launches)
    achieved = fast_bytes / (fast_ms / 1000.0) / 1e9 if fast_ms > 0 else 0.0
    b_frame = 2 * sp + 60 * (nkp / B)                                       # B_ext
    if wk.det:
        b_frame += 3.333 * W * H + 36 * (nmk / B)                           # B_aru
    if wk.matcher:
        b_frame += 36 * (nkp / B + 1000) + 4 * (nkp / B)                    # B_mat
    line = {
        "metric": wl["metric"], "value": fps, "unit": "frames/s", "n_gpus": ctx.world, "steps": steps, "warmup": warmup,
        "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": wl["name"], "frames_per_gpu": B * nsub, "l2": "flushed between steps (256 MiB fill, untimed)",
                   "timing": "CUDA events on the launching stream, one pair per step, max over ranks",
                   "streams": "extractor + matcher on the launching stream, detector on a second, higher-priority stream" + ("; sub-batches alternate between two sets of handles, result slots and streams" if len(wk.sets) > 1 else ""),
                   "collate": ("gather-to-root of the packed result slots through the library's own NCCL communicator (b200_collate_gather, grouped ncclSend/ncclRecv), two "
                               "transfers per %s inside the step: keypoints + descriptors + counts while the matcher runs, markers + matches after the join"
                               % ("sub-batch" if nsub > 1 else "step")) if ctx.world > 1 else "none (1 GPU)"},
        "e2e": {"value": frames / (e2e_ms / 1000.0) if do_e2e else None, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "b200_frontend_host (pinned host buffers, chunked H2D overlapped with compute)" +
                       ("; then b200_frontend_collate_host: every rank's results gathered over NCCL and downloaded into rank 0's host buffers" if ctx.world > 1 else "")},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "k_fast", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic("k_fast", wl, stage_frames), "peak_source": peak_src, "ms_per_launch": fast_ms,
                     "algorithmic_bytes_per_launch": int(fast_bytes), "frames_per_launch": int(stage_frames),
                     "extractor_stage_ms": {"pyramid": stage[0] / steps, "fast": stage[1] / steps, "quadtree": stage[2] / steps, "describe": stage[3] / steps},
                     "whole_step": {"algorithmic_bytes_per_frame": b_frame, "achieved_gbs": fps / ctx.world * b_frame / 1e9,
                                    "frac": fps / ctx.world * b_frame / 1e9 / peak}},
        "wall_s_timed_region": wall,
        "per_frame": {"keypoints": nkp / B, "markers": nmk / B, "matches": nmatch / B},
    }
    if ctx.world > 1:
        line["collated_parity"] = parity
        line["nvlink"] = {"bytes_received_by_rank0_per_step": int((traffic0[1] - traffic_before[1]) // max(1, steps + warmup)),
                          "bytes_sent_per_other_rank_per_step": int(wk.pack.total_bytes) * nsub,
                          "note": "payload of the grouped ncclSend/ncclRecv (b200_collate_traffic); only rank 0 receives"}
    if not headline:
        for k in ("higher_is_better", "scaling", "vs_baseline", "dtype", "data"):
            line.pop(k)
    return line


def c1_line(ctx):
    """C1 (BASELINE.json configs[0]): ONE 640x480 frame per call through the C++ adapters' ORBextractor::operator() + MarkerDetector::detect()
    (tools/c1_latency.cpp, the calls of src/Frame.cc:91,142), median latency of 200 calls, beside the 1-thread CPU reference on the same frames"""
    from orb_slam2_aruco_b200 import build as b, synth
    exe = b.C1_EXE
    if not os.path.exists(exe):
        b.build_tools()
    nfr = 8
    frames = synth.make_batch(nfr, 640, 480, 20, "ARUCO", first=500000)
    path = "/tmp/b200_c1_frames.raw"
    frames.tofile(path)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(ctx.local)))
    r = subprocess.run([exe, path, "640", "480", str(nfr), "ARUCO", "1000", "200", "20"], capture_output=True, text=True, timeout=300, env=env)
    if r.returncode != 0:
        return {"error": (r.stderr or r.stdout)[-300:]}
    d = json.loads(r.stdout.strip().splitlines()[-1])
    wl = dict(nfeatures=1000, markers=20, match=False, batch=nfr)
    cpu_reference_run(frames[:2], wl, 1, None, DICT="ARUCO")
    secs, kind, what = cpu_reference_run(frames, wl, 1, None, DICT="ARUCO")
    d.update({"workload": "C1: single 640x480 gray frame per call, 8 levels, 1000 ORB features, dictionary ARUCO (20 planted markers): ORB_SLAM2::ORBextractor::operator() + "
                          "aruco::MarkerDetector::detect(image, camParams, 0.187) through include/b200slam_adapters.hpp, host image in, std::vector<cv::KeyPoint> / cv::Mat / "
                          "std::vector<aruco::Marker> (with IPPE poses) out",
              "metric": "latency per frame (ms), median of 200 calls after 20 warm-up calls", "frames_per_s": 1000.0 / d["latency_ms_median"],
              "cpu_reference": {"latency_ms": 1000.0 * secs / nfr, "cores": 1, "kind": kind, "sample": "%d frames, 1 thread; %s" % (nfr, what)},
              "speedup_vs_cpu_reference_1_thread": (1000.0 * secs / nfr) / d["latency_ms_median"],
              "one_call": "one_call_ms_median = the same frame through ONE b200_frontend_host call (n = 1: extractor and detector concurrently on two streams) plus "
                          "b200_aruco_pose_host: what replacing the two calls of Frame::Frame by one gives; plain arrays out (no std::vector / cv::Mat assembly)"})
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-API leg")
    ap.add_argument("--no-extras", action="store_true", help="headline workload only (no extra_workloads)")
    ap.add_argument("--extras", default="C4,C5,C3R,C1", help="comma list of extra workloads measured after the headline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    import torch.distributed as dist
    from orb_slam2_aruco_b200 import shard

    ctx = Ctx()
    ctx.world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.rank = int(os.environ.get("RANK", "0"))
    ctx.local = int(os.environ.get("LOCAL_RANK", "0"))
    ctx.dist = dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(ctx.local)
    ctx.dev = torch.device("cuda", ctx.local)
    ctx.collator = None
    if ctx.world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=ctx.dev)       # plumbing only: rendezvous, barriers, max over ranks
        ctx.collator = shard.Collator.from_torch_distributed(ctx.local)      # the data path: the library's own NCCL communicator
    ctx.flush = torch.empty(256 << 20, dtype=torch.uint8, device=ctx.dev)    # > 126 MB L2

    sampler = ClockSampler(ctx.local)
    if ctx.rank == 0:
        sampler.start()
    line = workload_line(ctx, wl, args.steps, args.warmup, do_e2e=not args.no_e2e, headline=True)
    if ctx.rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)
        line["clocks"] = sampler.summary()
        if not args.no_cpu_baseline and ctx.world == 1:
            B, W, H = wl["batch"], wl["w"], wl["h"]
            sample = max(16, min(B, 256 * 640 * 480 // (W * H)))    # ~10 s of single-thread CPU work whatever the frame size
            ref_set = cpu_ref_set(wl) if wl["match"] else None
            secs, kind, what = cpu_reference_run(frames_for(wl, 0)[:sample], wl, 1, ref_set)
            line["cpu_baseline"] = {"value": sample / secs, "unit": "frames/s", "cores": 1, "kind": kind,
                                    "sample": "first %d of the %d frames, 1 thread; %s" % (sample, B, what)}
    extras = {}
    if not args.no_extras and args.workload == "C3":
        for name in [e for e in args.extras.split(",") if e]:
            try:
                if name == "C1":
                    if ctx.world == 1:
                        extras["C1"] = c1_line(ctx)
                    continue
                r = workload_line(ctx, WORKLOADS[name], 3, 3, do_e2e=not args.no_e2e)
                if ctx.rank == 0:
                    extras[name] = r
            except SystemExit:
                raise                                        # a collation that delivers wrong data prints nothing
            except Exception as e:                           # an extra must not take the headline down with it
                if ctx.world > 1:
                    raise
                extras[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
            torch.cuda.empty_cache()
    if ctx.rank == 0:
        if extras:
            line["extra_workloads"] = extras
        print(json.dumps(line))
    if ctx.world > 1:
        ctx.collator.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
