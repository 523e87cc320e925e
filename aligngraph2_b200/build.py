"""In-tree build of libag2_b200.so (hand-written CUDA for sm_100a + the C ABI of include/ag2_b200.h)."""
from __future__ import annotations

import os
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SO = os.path.join(PKG, "libag2_b200.so")
SOURCES = [os.path.join(PKG, "csrc", "ag2_b200.cu")]
HEADERS = [os.path.join(PKG, "csrc", n) for n in ("xdrop_device.cuh", "xdrop_lane.cuh", "seed_device.cuh", "index_kernels.cuh",
                                                   "rescue_device.cuh", "map_kernels.cuh")] + [ os.path.join(ROOT, "include", "ag2_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


HOST_SRC = os.path.join(PKG, "host", "mecat2ref_main.cpp")
HOST_BIN = os.path.join(PKG, "bin", "mecat2ref")


def build_host(force: bool = False) -> str:
    """The drop-in `mecat2ref` executable (C++ host over the C ABI)."""
    if not force and os.path.exists(HOST_BIN) and os.path.getmtime(HOST_BIN) > max(os.path.getmtime(HOST_SRC), os.path.getmtime(SO)):
        return HOST_BIN
    os.makedirs(os.path.dirname(HOST_BIN), exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-Wall", "-o", HOST_BIN, HOST_SRC, "-L" + PKG, "-lag2_b200",
                    "-Wl,-rpath,$ORIGIN/.."], check=True, cwd=ROOT)
    return HOST_BIN


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
