"""Writes tests/golden/match_ref.npz: seeded matcher inputs (tests/match_cases.py) and the answers of the reference's OWN src/ORBmatcher.cc,
compiled unmodified into oracle/_ref/libref_match.so (oracle/Makefile, needs /root/reference).  The file travels to the GPU box, where neither
the reference nor oracle/_ref needs to exist: tests/test_oracle_match_vs_ref.py replays it through oracle/match_oracle.cpp on the CPU and
tests/test_match_gpu.py through the CUDA path.

    python tests/golden/make_match_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import match_cases as mc
import match_cases2 as m2
import oracle

BOW_CONFIGS = [(0.6, 1), (0.75, 0), (0.9, 1), (1.5, 1)]
INIT_CONFIGS = [(100, 0.9, 1), (30, 0.9, 1), (10, 0.7, 0)]
POINTS_CONFIGS = [(1.0, 0.8), (3.0, 0.8), (5.0, 0.6)]
LAST_CONFIGS = [(7.0, 1), (15.0, 1), (15.0, 0)]
RELOC_CONFIGS = [(10.0, 100, 1), (3.0, 64, 1), (10.0, 100, 0)]
# match_ref2.npz: the KeyFrame-side members
TRI_CONFIGS = [0, 1]
FUSE_CONFIGS = [3.0, 6.0]
FUSE_SIM3_CONFIGS = [4.0, 10.0]
LOOP_CONFIGS = [10, 4]
SIM3_CONFIGS = [7.5, 3.0]


def main2(R):
    out = {}

    def put(prefix, d):
        for k, v in d.items():
            out[prefix + "." + k] = v

    c = m2.triangulation_inputs()
    put("tri", c)
    for j, ori in enumerate(TRI_CONFIGS):
        n, m = m2.run_triangulation(R, "ref", c, ori)
        put("tri.%d" % j, dict(cfg=np.array([ori]), n=np.int32(n), matches12=m))
    c = m2.keyframe_points_inputs()
    put("fuse", c)
    for j, th in enumerate(FUSE_CONFIGS):
        n, idx, act, qx, ql, qm = m2.run_fuse(R, "ref", c, th)
        put("fuse.%d" % j, dict(cfg=np.array([th]), n=np.int32(n), fused_idx=idx, action=act, q_xyr=qx, q_lev=ql, q_mp=qm))
    c = m2.keyframe_points_inputs(seed=54, sim3=True)
    put("scw", c)
    for j, (th, th2) in enumerate(zip(FUSE_SIM3_CONFIGS, LOOP_CONFIGS)):
        n, rep, add, qx, ql, qm = m2.run_fuse_sim3(R, "ref", c, th)
        n2, matched, qx2, ql2, qm2 = m2.run_loop(R, "ref", c, th2)
        put("scw.%d" % j, dict(cfg=np.array([th, th2]), n=np.int32(n), replace_idx=rep, added_idx=add, q_xyr=qx, q_lev=ql, q_mp=qm,
                               n_loop=np.int32(n2), matched=matched, q_xyr_loop=qx2, q_lev_loop=ql2, q_mp_loop=qm2))
    c = m2.sim3_inputs()
    put("sim3", c)
    for j, th in enumerate(SIM3_CONFIGS):
        n, m12, qx, ql, qm = m2.run_sim3(R, "ref", c, th)
        put("sim3.%d" % j, dict(cfg=np.array([th]), n=np.int32(n), matches12=m12, q_xyr=qx, q_lev=ql, q_mp=qm))
    path = os.path.join(HERE, "match_ref2.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", len(out), "arrays")


def main():
    R = oracle.ref_match()
    if R is None:
        raise SystemExit("oracle/_ref/libref_match.so not built (needs /root/reference): make -C oracle ref")
    mc.NFEATURES = 500
    main2(R)
    out = {}

    def put(prefix, d):
        for k, v in d.items():
            out[prefix + "." + k] = v

    c = mc.bow_inputs()
    put("bow", c)
    for j, (ratio, ori) in enumerate(BOW_CONFIGS):
        n, m = mc.run_bow(R, "ref", c, ratio, ori)
        n2, m2 = mc.run_bow_kfkf(R, "ref", c, ratio, ori)
        put("bow.%d" % j, dict(cfg=np.array([ratio, ori]), n=np.int32(n), matches=m, n_kfkf=np.int32(n2), matches12=m2))
    c1 = mc.one_node(c)                                       # brute force = one node, every MapPoint good (the benched configuration)
    for j, (ratio, ori) in enumerate(BOW_CONFIGS):
        n, m = mc.run_bow(R, "ref", c1, ratio, ori)
        n2, m2 = mc.run_bow_kfkf(R, "ref", c1, ratio, ori)
        put("bf.%d" % j, dict(cfg=np.array([ratio, ori]), n=np.int32(n), matches=m, n_kfkf=np.int32(n2), matches12=m2))
    c = mc.init_inputs()
    put("init", c)
    for j, (window, ratio, ori) in enumerate(INIT_CONFIGS):
        prev = c["prev"]
        for rep in range(2):                                  # the second call continues from the updated vbPrevMatched
            n, m, prev = mc.run_init(R, "ref", c, prev, window, ratio, ori)
            put("init.%d.%d" % (j, rep), dict(cfg=np.array([window, ratio, ori]), n=np.int32(n), matches12=m, prev=prev))
    for kind, make, run, configs in (("points", mc.points_inputs, mc.ref_points, POINTS_CONFIGS), ("last", mc.last_inputs, mc.ref_last, LAST_CONFIGS),
                                     ("reloc", mc.reloc_inputs, mc.ref_reloc, RELOC_CONFIGS)):
        c = make()
        put(kind, c)
        for j, cfg in enumerate(configs):
            n, assign, qx, ql, qm = run(R, c, *cfg)
            put("%s.%d" % (kind, j), dict(cfg=np.array(cfg, np.float64), n=np.int32(n), assign=assign, q_xyr=qx, q_lev=ql, q_mp=qm))
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (64, 32)).astype(np.uint8); b = rng.integers(0, 256, (64, 32)).astype(np.uint8)
    b[:8] = a[:8]; b[8:16] = ~a[8:16]
    out["dist.a"] = a; out["dist.b"] = b
    out["dist.d"] = np.array([R.ref_descriptor_distance(mc.P(a[i]), mc.P(b[i])) for i in range(64)], np.int32)
    path = os.path.join(HERE, "match_ref.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
