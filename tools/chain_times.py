#!/usr/bin/env python
"""Which chain bounds the C3 step: the detector alone, extractor + matcher alone, and both concurrently (the bench's step), device-timed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

ctx = bench.Ctx()
ctx.dev, ctx.local, ctx.rank, ctx.world = torch.device("cuda", 0), 0, 0, 1
wk = bench.Workload(ctx, bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C3"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=ctx.dev)
v, imgs = wk.v, wk.d_imgs[:wk.B]


def timed(fn, n=10):
    ts = []
    for i in range(n + 3):
        flush.fill_(i & 255)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        with torch.cuda.stream(wk.s_main):
            e0.record(wk.s_main)
            fn()
            e1.record(wk.s_main)
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)


def det_only():
    wk.det.detect_batch_device(imgs, v["markers"], v["marker_counts"], wk.s_main)


def orb_only():
    wk.ex.extract_batch_device(imgs, v["kps"], v["desc"], v["counts"], wk.s_main)
    wk.matcher.SearchByBoW_device(wk.d_rdesc, wk.d_rkps, wk.n_ref, v["desc"], v["kps"], v["counts"], v["matches"], v["n_matches"], wk.s_main)


print("detector alone      %.3f ms" % timed(det_only))
print("extractor + matcher %.3f ms" % timed(orb_only))
print("both (bench step)   %.3f ms" % timed(wk.step))
