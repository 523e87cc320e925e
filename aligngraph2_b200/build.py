"""In-tree build of libag2_b200.so (hand-written CUDA for sm_100a + the C ABI of include/ag2_b200.h)."""
from __future__ import annotations

import os
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SO = os.path.join(PKG, "libag2_b200.so")
SOURCES = [os.path.join(PKG, "csrc", "ag2_b200.cu")]
HEADERS = [os.path.join(PKG, "csrc", "xdrop_device.cuh"), os.path.join(PKG, "csrc", "xdrop_lane.cuh"), os.path.join(ROOT, "include", "ag2_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + SOURCES
    subprocess.run(cmd, check=True, cwd=ROOT)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
