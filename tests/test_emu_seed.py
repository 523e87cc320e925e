"""The product's seeding / candidate device code (aligngraph2_b200/csrc/seed_device.cuh, compiled unchanged with
-DAG2_EMU by tests/emu/build.sh) on the CPU against the oracle, both passes, on the stress fixture.  The index
arrays are the oracle's (the kernels that build them are checked on the GPU in tests/test_gpu_index.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

from conftest import GOLDEN

EMU_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")


def test_emulated_seeding_matches_oracle():
    from oracle.binding import IndexOracle
    subprocess.run([os.path.join(EMU_DIR, "build.sh")], check=True)
    E = C.CDLL(os.path.join(EMU_DIR, "libemu_seed.so"))
    E.emu_seed_batch.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_void_p,
                                 C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    z = np.load(os.path.join(GOLDEN, "mapper_stress.npz"))
    bases, offs = z["bases"].tobytes(), z["offsets"].astype(np.int64)
    io = IndexOracle(z["genome"].tobytes(), bases, offs)
    n, maxc = len(offs) - 1, 10
    total = 0
    for p in (0, 1):
        out = np.zeros(n * maxc * 10, dtype=np.int64)
        nc = np.zeros(n, dtype=np.int32)
        E.emu_seed_batch(len(io.ref), io.cnt.ctypes.data, io.off.ctypes.data, io.pos.ctypes.data, io.vote.ctypes.data, 200,
                         bases, offs.ctypes.data, n, p, maxc, out.ctypes.data, nc.ctypes.data)
        for r in range(n):
            exp = io.candidates(bases[offs[r]:offs[r + 1]], p)
            got = [tuple(int(v) for v in out[(r * maxc + i) * 10:(r * maxc + i) * 10 + 10]) for i in range(nc[r])]
            assert got == exp, (p, r)
            total += len(exp)
    assert total > 500


def test_emulated_cta_seeding_matches_oracle():
    """The CTA path (seed_cta.cuh: events -> radix sort -> runs; heavy blocks through insert_loc; candidate scan by a
    warp) on the same fixture: every read either gives the oracle's candidate list or is reported as overflow, and with
    a roomy `cap` nothing overflows."""
    from oracle.binding import IndexOracle
    subprocess.run([os.path.join(EMU_DIR, "build.sh")], check=True)
    E = C.CDLL(os.path.join(EMU_DIR, "libemu_seed.so"))
    E.emu_seed_cta_batch.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_void_p,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    z = np.load(os.path.join(GOLDEN, "mapper_stress.npz"))
    bases, offs = z["bases"].tobytes(), z["offsets"].astype(np.int64)
    io = IndexOracle(z["genome"].tobytes(), bases, offs)
    n, maxc = len(offs) - 1, 10
    for cap, may_overflow in ((32768, False), (1024, True), (256, True)):
        total = 0
        for p in (0, 1):
            out = np.zeros(n * maxc * 10, dtype=np.int64)
            nc = np.zeros(n, dtype=np.int32)
            novf = E.emu_seed_cta_batch(len(io.ref), io.cnt.ctypes.data, io.off.ctypes.data, io.pos.ctypes.data, io.vote.ctypes.data, 200,
                                        bases, offs.ctypes.data, n, p, maxc, cap, out.ctypes.data, nc.ctypes.data)
            assert novf == int((nc < 0).sum())
            assert may_overflow or novf == 0
            for r in range(n):
                if nc[r] < 0:
                    continue
                exp = io.candidates(bases[offs[r]:offs[r + 1]], p)
                got = [tuple(int(v) for v in out[(r * maxc + i) * 10:(r * maxc + i) * 10 + 10]) for i in range(nc[r])]
                assert got == exp, (cap, p, r)
                total += len(exp)
        assert total > (500 if not may_overflow else 0)
