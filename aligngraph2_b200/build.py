"""In-tree build of libag2_b200.so (hand-written CUDA for sm_100a + the C ABI of include/ag2_b200.h)."""
from __future__ import annotations

import os
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SO = os.path.join(PKG, "libag2_b200.so")
SOURCES = [os.path.join(PKG, "csrc", "ag2_b200.cu"), os.path.join(PKG, "csrc", "pagraph.cu"), os.path.join(PKG, "host", "pg_job.cpp"),
           os.path.join(PKG, "host", "pg_travel.cpp")]
HEADERS = [os.path.join(PKG, "csrc", n) for n in ("xdrop_device.cuh", "xdrop_lane.cuh", "seed_device.cuh", "index_kernels.cuh",
                                                   "rescue_device.cuh", "seed_cta.cuh", "plan_cta.cuh", "map_kernels.cuh", "kmer_kernels.cuh", "pagraph_kernels.cuh", "xdrop_pair.cuh", "h2ops.cuh")] + [
    os.path.join(ROOT, "include", "ag2_b200.h"), os.path.join(ROOT, "include", "ag2_pagraph.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


HOST_SRC = os.path.join(PKG, "host", "mecat2ref_main.cpp")
HOST_BIN = os.path.join(PKG, "bin", "mecat2ref")
HOST_PROGRAMS = {"mecat2ref": "mecat2ref_main.cpp", "kmer_counter": "kmer_counter_main.cpp", "pagraph": "pagraph_main.cpp"}


def build_host(force: bool = False, name: str = "mecat2ref") -> str:
    """The drop-in executables (C++ hosts over the C ABI): `mecat2ref`, `kmer_counter`, `pagraph`.
    Builds all of them, returns the path of `name`."""
    os.makedirs(os.path.dirname(HOST_BIN), exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    for prog, src in HOST_PROGRAMS.items():
        src_path, out = os.path.join(PKG, "host", src), os.path.join(PKG, "bin", prog)
        deps = [src_path, SO, os.path.join(PKG, "host", "shard_split.h")]
        if not force and os.path.exists(out) and os.path.getmtime(out) > max(os.path.getmtime(p) for p in deps):
            continue
        subprocess.run([cxx, "-O2", "-std=c++17", "-Wall", "-pthread", "-o", out, src_path, "-L" + PKG, "-lag2_b200",
                        "-Wl,-rpath,$ORIGIN/.."], check=True, cwd=ROOT)
    return os.path.join(PKG, "bin", name)


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
