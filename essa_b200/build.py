"""Builds essa_b200/lib/libessa_b200.so (CUDA kernels + C ABI + C++ World facade) for sm_100a.

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the
repo snapshot.  Usage:  python -m essa_b200.build [--force]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(OUT_DIR, "obj")
LIB = os.path.join(OUT_DIR, "libessa_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CUFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
           "--fmad=true"] + os.environ.get("ESSA_NVCC_EXTRA", "").split()  # e.g. -DESSA_STAGE_CLOCKS (diagnostics)
CXX = os.environ.get("CXX", "g++")
HOSTFLAGS = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-ffp-contract=off", "-I/usr/local/cuda/include"]


def sources():
    cu = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    cpp = sorted(f for f in os.listdir(CSRC) if f.endswith(".cpp"))
    return cu, cpp


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".hpp", ".h", ".cpp")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return max(m, os.path.getmtime(os.path.abspath(__file__)))


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError("nvcc not found at %s (set NVCC)" % NVCC)
    cu, cpp = sources()
    jobs = []
    for f in cu:
        obj = os.path.join(OBJ_DIR, f + ".o")
        jobs.append(([NVCC, *ARCH, *CUFLAGS, "-Xptxas", "-v", "-c", os.path.join(CSRC, f), "-o", obj], obj))
    for f in cpp:
        obj = os.path.join(OBJ_DIR, f + ".o")
        jobs.append(([CXX, *HOSTFLAGS, "-c", os.path.join(CSRC, f), "-o", obj], obj))

    def run(job):
        cmd, obj = job
        p = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, obj, p

    logs = []
    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        for cmd, obj, p in ex.map(run, jobs):
            logs.append("$ " + " ".join(cmd) + "\n" + p.stdout + p.stderr)
            if p.returncode != 0:
                raise RuntimeError("compile failed:\n" + logs[-1])
    with open(os.path.join(OUT_DIR, "build.log"), "w") as f:
        f.write("\n".join(logs))
    link = [NVCC, *ARCH, "-shared", "-o", LIB, *[j[1] for j in jobs], "-ldl", "-cudart", "static"]
    p = subprocess.run(link, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("link failed:\n" + p.stdout + p.stderr)
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
