#!/usr/bin/env python
"""bench.py -- frames/s of the B200 front end on BASELINE.json's workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C2|C3]

A "step" is one pass of the hot path over one batch of synthetic frames (config C2 of BASELINE.json:
batch=256 synthetic 640x480 frames, 1000 features, 8 levels; C3 adds ArUco + brute-force matching once built).
`value` is measured with the batch already resident in HBM (CUDA events on the launching stream, one event pair
per step, L2 flushed between steps); `e2e` goes through the reference-facing host-pointer C-ABI call with
pinned host buffers, H2D and D2H inside the timed region.  Multi-GPU (torchrun, one rank per GPU): frames are
independent, every rank processes its own batch (weak scaling), results are collated on rank 0 with one NCCL
all_gather of fixed-size slots inside the timed region; max over ranks.

--impl reference times the reference's own CPU extractor (oracle/_ref: src/ORBextractor.cc compiled unmodified
on the cv shim; falls back to the oracle port) with all host threads, on the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    "C2": dict(name="C2: batch=256 synthetic 640x480 gray frames, 1000 ORB features, 8 levels x1.2, FAST 20/7, extract-only",
               batch=256, w=640, h=480, nfeatures=1000, markers=0),
}
SIGMA_P = {(640, 480): 950532}      # pixels over the 8 levels


def frames_for(wl, rank, batch=None):
    """deterministic synthetic frames; cached under /tmp because numpy generation takes ~50 ms per frame"""
    from orb_slam2_aruco_b200 import synth
    n = batch or wl["batch"]
    path = "/tmp/b200_frames_%dx%d_m%d_r%d_n%d.npy" % (wl["w"], wl["h"], wl["markers"], rank, n)
    if os.path.exists(path):
        try:
            return np.load(path)
        except Exception:
            pass
    imgs = synth.make_batch(n, wl["w"], wl["h"], wl["markers"], first=rank * 100000)
    try:
        np.save(path, imgs)
    except Exception:
        pass
    return imgs


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.stop_flag, self.rows = gpu, False, []

    def run(self):
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
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_run(imgs, wl, nthreads):
    """the reference's own extractor (oracle/_ref) or the oracle port on `imgs`; returns (seconds, kind)"""
    import ctypes as C
    import oracle
    n, h, w = imgs.shape
    cap = wl["nfeatures"] + 200
    r = oracle.ref()
    t0 = time.perf_counter()
    if r is not None:
        kps = np.zeros((n, cap, 7), np.float32); desc = np.zeros((n, cap, 32), np.uint8); cnt = np.zeros(n, np.int32)
        r.ref_orb_extract_batch(imgs.ctypes.data_as(C.c_void_p), n, w, h, w, C.c_long(w * h), wl["nfeatures"], C.c_float(1.2), 8, 20, 7,
                                kps.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), cap, nthreads)
        kind = "reference"
    else:
        oracle.orb_extract_batch(imgs, wl["nfeatures"], nthreads=nthreads)
        kind = "port"
    return time.perf_counter() - t0, kind


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    try:
        ncores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    sample = min(wl["batch"], max(32, 2 * ncores))          # bounded sample of the workload per step
    imgs = frames_for(wl, 0)[:sample]
    for _ in range(args.warmup):
        cpu_reference_run(imgs, wl, ncores)
    t = 0.0
    kind = "port"
    for _ in range(args.steps):
        dt, kind = cpu_reference_run(imgs, wl, ncores)
        t += dt
    fps = sample * args.steps / t
    line = {"impl": "reference", "metric": "frames/sec (extract)", "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": {"workload": wl["name"]},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": ncores, "kind": kind,
                             "sample": "%d of the %d frames per step, %d host threads" % (sample, wl["batch"], ncores)},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    import torch.distributed as dist
    from orb_slam2_aruco_b200 import _lib
    from orb_slam2_aruco_b200.api import ORBextractor

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
    cap = ex.cap
    d_imgs = torch.from_numpy(imgs_np).to(dev)
    d_kps = torch.zeros((B, cap, 7), dtype=torch.float32, device=dev)
    d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev)
    d_counts = torch.zeros((B,), dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    # collation on rank 0: fixed-size slots, one all_gather per step
    gather = None
    if world > 1:
        gather = [torch.empty_like(d_desc) for _ in range(world)], [torch.empty_like(d_kps) for _ in range(world)], \
                 [torch.empty_like(d_counts) for _ in range(world)]

    def step():
        ex.extract_batch_device(d_imgs, d_kps, d_desc, d_counts, stream)
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_gather(gather[0], d_desc)
                dist.all_gather(gather[1], d_kps)
                dist.all_gather(gather[2], d_counts)

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
        with torch.cuda.stream(stream):
            flush.fill_(i & 0xff)          # evict the batch from L2 (not timed)
            ev[i][0].record(stream)
        step()
        with torch.cuda.stream(stream):
            ev[i][1].record(stream)
        stream.synchronize()
        stage += ex.stage_ms()
    sync_all()
    t_wall = time.perf_counter() - t_wall0
    launches = _lib.launch_count() - launches0
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(ms_steps))
    ex.set_profile(False)

    # ---- end to end through the reference-facing host-pointer C-ABI (pinned buffers) ----------------
    h_imgs = torch.from_numpy(imgs_np).pin_memory()
    h_kps = torch.zeros((B, cap, 7), dtype=torch.float32).pin_memory()
    h_desc = torch.zeros((B, cap, 32), dtype=torch.uint8).pin_memory()
    h_counts = torch.zeros((B,), dtype=torch.int32).pin_memory()
    out = (h_kps.numpy().view(_lib.KP_DTYPE).reshape(B, cap), h_desc.numpy(), h_counts.numpy())
    for _ in range(2):
        ex.extract_batch(h_imgs.numpy(), out=out)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ex.extract_batch(h_imgs.numpy(), out=out)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    nkp = int(h_counts.sum())
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
        ncand_bytes = 4 * 10000                                   # ~10k candidates x 4 B per frame (measured on this workload)
        fast_bytes = (sp + ncand_bytes) * B                       # k_fast: every level pixel read once + candidate slots written
        fast_ms = stage[1] / args.steps
        achieved = fast_bytes / (fast_ms / 1000.0) / 1e9
        b_ext = 2 * sp + 60 * (nkp / B)                           # whole-extractor algorithmic bytes / frame
        line = {
            "metric": "frames/sec (extract)", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": wl["name"], "frames_per_gpu": B, "l2": "flushed between steps (256 MiB fill, untimed)",
                       "timing": "CUDA events on the launching stream, one pair per step, max over ranks",
                       "collate": "nccl all_gather of fixed slots inside the step" if world > 1 else "none (1 GPU)"},
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": int(B * W * H),
                    "d2h_bytes_per_step": int(B * cap * 60 + B * 4), "api": "b200_orb_extract_host (pinned host buffers, chunked H2D overlap)"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "k_fast", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "ms_per_launch": fast_ms,
                         "algorithmic_bytes_per_launch": int(fast_bytes),
                         "stage_ms": {"pyramid": stage[0] / args.steps, "fast": stage[1] / args.steps, "quadtree": stage[2] / args.steps,
                                      "describe": stage[3] / args.steps},
                         "whole_step": {"algorithmic_bytes_per_frame": b_ext, "achieved_gbs": fps / world * b_ext / 1e9,
                                        "frac": fps / world * b_ext / 1e9 / peak}},
            "clocks": sampler.summary(),
            "wall_s_timed_region": t_wall,
            "keypoints_per_frame": nkp / B,
        }
        if not args.no_cpu_baseline and world == 1:
            sample = 128
            secs, kind = cpu_reference_run(imgs_np[:sample], wl, 1)
            line["cpu_baseline"] = {"value": sample / secs, "unit": "frames/s", "cores": 1, "kind": kind,
                                    "sample": "first %d of the %d frames, 1 thread" % (sample, B)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
