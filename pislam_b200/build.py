"""Builds pislam_b200/lib/libpislam_b200.so (the C-ABI library, CUDA kernels for sm_100a) in-tree.

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels with the tree.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libpislam_b200.so")
SOURCES = ["pislam_b200.cu"]
HEADERS = ["device_ops.cuh", "front_kernels.cuh", "orb_kernels.cuh", "pyramid_kernels.cuh", "match_kernels.cuh", "brief_pattern.h",
           os.path.join("..", "..", "include", "pislam_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared"]


def _nvcc():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.sep not in p or os.path.exists(p)):
            return p
    raise RuntimeError("nvcc not found")


def stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_library(force=False, verbose=False, out=None, extra=()):
    """`out` / `extra` build a variant for A/B runs (tools/gpu_front_ab.py), e.g. extra=["-DPB_FRONT_MIN_CTAS=8"]."""
    if out is None and not force and not stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", out or LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout)
    if verbose:
        print(r.stdout)
    return out or LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose="-v" in sys.argv))
