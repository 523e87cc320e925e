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
    L.emu_lane_batch.argtypes = [C.c_char_p, C.c_long, C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 5
    L.emu_pair_batch.argtypes = L.emu_lane_batch.argtypes
    return L


def emu_lane_batch(L, ref, reads, cands, pair=False):
    """The lane path (xdrop_lane.cuh) for a batch: setup -> lane kernel body (one emulated warp, queue
    refill) -> wide rerun -> finalize -> assemble.  pair=True: the pair path (xdrop_pair.cuh) in front of it;
    the third return value is then (wide, handed to the lane kernel)."""
    offs = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in reads], out=offs[1:])
    cat = b"".join(reads)
    n = len(cands)
    c = np.ascontiguousarray(np.array(cands, dtype=np.int64))
    rec = np.zeros(8 * n, dtype=np.int64)
    aoff = np.zeros(n, dtype=np.int64)
    cap = int(offs[-1]) * 3 + 100000
    qa, ta = C.create_string_buffer(cap), C.create_string_buffer(cap)
    st = np.zeros(3, dtype=np.int64)
    (L.emu_pair_batch if pair else L.emu_lane_batch)(ref, len(ref), cat, offs.ctypes.data, len(reads), c.ctypes.data, n, rec.ctypes.data,
                     aoff.ctypes.data, qa, ta, st.ctypes.data)
    out = []
    for i in range(n):
        r = rec[8 * i:8 * i + 8]
        ln, a = int(r[5]), int(aoff[i])
        out.append(dict(ok=int(r[0]), qb=int(r[1]), qe=int(r[2]), sb=int(r[3]), se=int(r[4]), aln_size=ln,
                        qaln=qa.raw[a:a + ln] if r[0] else b"", taln=ta.raw[a:a + ln] if r[0] else b"",
                        modes=(int(r[6]), int(r[7]))))
    return out, int(st[0]), ((int(st[1]), int(st[2])) if pair else int(st[1]))


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


def _stress_batch(oracle, seed, n):
    # more directions than lanes (queue refill), ragged reads, both strands, soft-masked / N bases,
    # seeds at the read ends, and unrelated extensions whose band leaves the lane window (wide rerun)
    rng = np.random.default_rng(seed)
    ref = synth.make_reference(rng, 40_000)
    reads, cands, exp, cells = [], [], [], 0
    for i in range(n):
        tl = int(rng.integers(60, 2600))
        rd, start, ops = synth.make_read(rng, ref, tl, False)
        cnt = np.ones(tl, dtype=np.int64)
        cnt[ops == 1] = 2
        cnt[ops == 2] = 0
        off = np.concatenate(([0], np.cumsum(cnt)))
        good = [p for p in range(0, max(1, tl - 13)) if (ops[p:p + 13] == 0).all()]
        p = good[int(rng.integers(0, len(good)))] if good else 0
        loc2 = int(off[p + 1] - 1) if good else 0
        loc1 = start + p + 1
        fwd = rd.tobytes()
        if i % 5 == 0:
            b = bytearray(fwd)
            for k in rng.integers(0, len(b), size=5):
                b[k] = ord("N")
            for k in rng.integers(0, len(b), size=20):
                b[k] |= 0x20
            fwd = bytes(b)
        strand = i & 1
        given = fwd if not strand else synth.orient(fwd, 1)
        if i % 11 == 3:
            loc2 = 0
        if i % 11 == 7:
            loc2 = len(fwd)
        reads.append(given)
        cands.append((i, strand, loc1, loc2))
        a = oracle.extend(ref.tobytes(), synth.orient(given, strand), loc1, loc2)
        exp.append(a)
        cells += a["cells"]
    return ref, reads, cands, exp, cells


def _check(exp, got):
    for a, b in zip(exp, got):
        assert a["ok"] == b["ok"]
        if a["ok"]:
            for k in ("qb", "qe", "sb", "se", "aln_size", "qaln", "taln"):
                assert a[k] == b[k], k


def test_emulated_lane_path_matches_oracle(emu, oracle):
    ref, reads, cands, exp, cells = _stress_batch(oracle, 21, 40)
    got, got_cells, wide = emu_lane_batch(emu, ref.tobytes(), reads, cands)
    _check(exp, got)
    assert got_cells == cells
    assert wide > 0, "no direction was handed to the wide path: that hand-over is part of this test"


@pytest.mark.parametrize("defer", [0])
def test_emulated_pair_path_matches_oracle(emu, oracle, defer):
    # the pair path (two directions per lane, packed 16-bit DP) in front of the lane path: same stress batch, more
    # directions than the 64 the emulated warp holds (refill), target blocks shorter than 32 and unrelated extensions
    # (handed to the lane kernel, then to the wide kernel)
    # defer = 1: directions whose last block is longer than 500 rows are set aside and run at the end of the launch
    ref, reads, cands, exp, cells = _stress_batch(oracle, 22, 80)
    emu.emu_set_defer(defer)
    try:
        got, got_cells, (wide, handed) = emu_lane_batch(emu, ref.tobytes(), reads, cands, pair=True)
    finally:
        emu.emu_set_defer(0)
    _check(exp, got)
    assert got_cells == cells
    assert handed > 0 and handed < len(cands), "the pair kernel must keep most directions and hand some over"


@pytest.mark.parametrize("defer", [0, 1])
def test_emulated_pair_path_long_reads(emu, oracle, defer):
    # full-length blocks (500 x 500), window moves, several blocks per direction
    d = synth.make_batch_torch(77, 150_000, 24, 4000)
    ref, bases, off = d["ref"].numpy(), d["bases"].numpy(), d["offsets"].numpy()
    reads = [bases[off[i]:off[i + 1]].tobytes() for i in range(24)]
    cands = [(i, int(d["strand"][i]), int(d["loc1"][i]), int(d["loc2"][i])) for i in range(24)]
    exp, cells = [], 0
    for i in range(24):
        a = oracle.extend(ref.tobytes(), synth.orient(reads[i], cands[i][1]), cands[i][2], cands[i][3])
        exp.append(a)
        cells += a["cells"]
    emu.emu_set_defer(defer)
    try:
        got, got_cells, (wide, handed) = emu_lane_batch(emu, ref.tobytes(), reads, cands, pair=True)
    finally:
        emu.emu_set_defer(0)
    _check(exp, got)
    assert got_cells == cells


def test_emulated_hand_overs_continued_by_the_row_parallel_form(emu, oracle):
    """What the consumer kernel does with a direction the pair kernel hands over (xdrop_stream_kernel ->
    continue_handed_over -> run_chain_resumed): continue it at the failing block on the row-parallel DP, writing the pair
    path's per-block segments.  Two batches: the stress batch (short target blocks, unrelated extensions: hand-overs at the
    first blocks) and reads with a tandem repeat deep inside (hand-overs after several blocks)."""
    emu.emu_set_row_resume(1)
    try:
        ref, reads, cands, exp, cells = _stress_batch(oracle, 22, 40)
        got, got_cells, (wide, handed) = emu_lane_batch(emu, ref.tobytes(), reads, cands, pair=True)
        _check(exp, got)
        assert got_cells == cells and handed > 0

        rng = np.random.default_rng(11)
        unit = synth.make_reference(rng, 7)
        ref = np.concatenate([synth.make_reference(rng, 5000), np.tile(unit, 400), synth.make_reference(rng, 5000)])
        reads, cands, exp, cells = [], [], [], 0
        for i in range(4):
            start = int(rng.integers(500, 1500))
            rd, _, _ = synth.make_read(rng, ref[start:start + 5001], 5000, False)
            strand = i & 1
            given = rd.tobytes() if not strand else synth.orient(rd.tobytes(), 1)
            reads.append(given)
            cands.append((i, strand, start + 301, 300))
            a = oracle.extend(ref.tobytes(), synth.orient(given, strand), start + 301, 300)
            exp.append(a)
            cells += a["cells"]
        got, got_cells, (wide, handed) = emu_lane_batch(emu, ref.tobytes(), reads, cands, pair=True)
        _check(exp, got)
        assert got_cells == cells and handed > 0
    finally:
        emu.emu_set_row_resume(0)
