"""Compile the CUDA sources of this package for sm_100a into orb_slam2_aruco_b200/libb200slam.so (in-tree)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200slam.so")
SOURCES = ["common.cu", "orb.cu", "match.cu", "aruco.cu", "pose.cu", "frame.cu", "bow.cu", "collate.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--fmad=false"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    if not os.path.exists(C1_EXE):
        return True
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "b200slam.h"),
                                                                 os.path.join(HERE, "..", "include", "b200slam_adapters.hpp"),
                                                                 os.path.join(HERE, "..", "tools", "c1_latency.cpp")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in SOURCES:
        o = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        objs.append(o)
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            print(" ".join(cmd)); print(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    build_tools()
    return LIB


C1_EXE = os.path.join(HERE, "c1_latency.bin")


def build_tools():
    """host programs over the C++ adapters (g++, header only): the C1 single-frame latency driver bench.py runs"""
    root = os.path.join(HERE, "..")
    cmd = ["g++", "-std=c++14", "-O2", "-I", os.path.join(root, "include"), os.path.join(root, "tools", "c1_latency.cpp"), "-o", C1_EXE,
           "-L", HERE, "-l:libb200slam.so", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return C1_EXE


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
