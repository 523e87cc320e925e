"""Runs the product's device code (aligngraph2_b200/csrc/xdrop_device.cuh, compiled unchanged with
-DAG2_EMU against tests/emu/warp_emu.h -- a lock-step one-warp emulation) and checks it against
the oracle.  This is how the kernel logic is exercised in the GPU-less build container; the real
parity tests are tests/test_gpu_extend.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from aligngraph2_b200 import synth
from conftest import load_npz_rows, random_block

EMU_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")


@pytest.fixture(scope="module")
def emu():
    subprocess.run([os.path.join(EMU_DIR, "build.sh")], check=True)
    L = C.CDLL(os.path.join(EMU_DIR, "libemu_xdrop.so"))
    L.emu_dp_block.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 7
    L.emu_extend.argtypes = [C.c_int, C.c_char_p, C.c_long, C.c_char_p, C.c_int, C.c_int, C.c_long, C.c_int,
                             C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def emu_block(L, K, A, B):
    A = np.ascontiguousarray(A, dtype=np.uint8)
    B = np.ascontiguousarray(B, dtype=np.uint8)
    ops = np.zeros(len(A) + len(B) + 8, dtype=np.uint8)
    ae, be, nops, ov = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    cells, inter = C.c_long(), C.c_long()
    L.emu_dp_block(K, A.ctypes.data, len(A), B.ctypes.data, len(B), C.byref(ae), C.byref(be), ops.ctypes.data,
                   C.byref(nops), C.byref(cells), C.byref(ov), C.byref(inter))
    return ae.value, be.value, ops[:nops.value].copy(), cells.value, ov.value, inter.value


def emu_extend(L, K, ref, read, strand, loc1, loc2):
    rec = (C.c_long * 8)()
    cap = len(read) * 3 + 4096
    qa, ta = C.create_string_buffer(cap), C.create_string_buffer(cap)
    ok = L.emu_extend(K, ref, len(ref), read, len(read), strand, loc1, loc2, rec, qa, ta)
    n = rec[4]
    return dict(ok=ok, qb=rec[0], qe=rec[1], sb=rec[2], se=rec[3], aln_size=n, cells=rec[5], wide=rec[6],
                interior=rec[7], qaln=qa.raw[:n], taln=ta.raw[:n])


def test_emulated_blocks_match_oracle(emu, oracle):
    rng = np.random.default_rng(2)
    interior = 0
    overflow4 = 0
    for _ in range(25):
        A, B = random_block(rng)
        s, ae, be, ops, cells = oracle.block(A, len(A), B, len(B))
        for K in (4, 23, 3):
            r = emu_block(emu, K, A, B)
            if r[4]:
                assert K != 23, "the 736-column kernel can never overflow"
                overflow4 += K == 4
                continue
            assert (r[0], r[1], r[3]) == (ae, be, cells)
            assert np.array_equal(r[2], ops)
            interior += r[5]
    assert interior > 0, "the interior-pruned-cell fix-up was never exercised"


def test_emulated_blocks_match_golden(emu):
    _, rows = load_npz_rows("xdrop_blocks.npz")
    for r in rows[:12]:
        A, B = r["A"], r["B"]
        if not bool(r["fwd"]):      # the device code always sees blocks in extension order
            A, B = A[::-1].copy(), B[::-1].copy()
        e = emu_block(emu, 23, A, B)
        assert (e[0], e[1]) == (int(r["ae"]), int(r["be"]))
        assert np.array_equal(e[2], r["ops"])


def test_emulated_extend_matches_golden(emu):
    z, rows = load_npz_rows("extend_candidates.npz")
    ref = z["ref"].tobytes()
    bases, off = z["bases"], z["offsets"]
    for i in (0, 1, 2, 5, 7):
        r = rows[i]
        e = emu_extend(emu, 4, ref, bases[off[i]:off[i + 1]].tobytes(), int(r["strand"]), int(r["loc1"]), int(r["loc2"]))
        assert e["ok"] == int(r["ok"])
        if e["ok"]:
            assert (e["qb"], e["qe"], e["sb"], e["se"]) == (int(r["qb"]), int(r["qe"]), int(r["sb"]), int(r["se"]))
            assert e["qaln"] == r["qaln"].tobytes() and e["taln"] == r["taln"].tobytes()
