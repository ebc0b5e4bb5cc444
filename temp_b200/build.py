"""In-tree build of libtemp_b200.so (hand-written sm_100a CUDA + the C ABI of include/temp_b200.h).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the tree.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libtemp_b200.so")
SOURCES = [os.path.join(HERE, "csrc", f) for f in ("temp_kernels.cu", "tc_kernels.cu", "tc_scan2.cu", "tc_wide.cu", "planner.cpp")]
HEADERS = [os.path.join(ROOT, "include", "temp_b200.h"), os.path.join(HERE, "csrc", "internal.h"),
           os.path.join(HERE, "csrc", "tc_common.cuh")]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    mt = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > mt for p in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not is_stale():
        return LIB_PATH
    cmd = [nvcc_path(), "--threads", "4", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-o", LIB_PATH] + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr))
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
