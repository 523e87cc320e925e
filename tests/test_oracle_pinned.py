"""Pins oracle/ag2_oracle.c: against the golden vectors produced by the unmodified reference
(tests/golden/gen_golden.py) and, where oracle/_ref exists, against the reference itself on
fresh random inputs."""
import numpy as np

from aligngraph2_b200 import synth
from conftest import load_npz_rows, random_block


def test_blocks_match_golden(oracle):
    _, rows = load_npz_rows("xdrop_blocks.npz")
    assert len(rows) >= 40
    for r in rows:
        A, B, fwd = r["A"], r["B"], bool(r["fwd"])
        s, ae, be, ops, cells = oracle.block(A, len(A), B, len(B), fwd)
        assert (s, ae, be) == (int(r["score"]), int(r["ae"]), int(r["be"]))
        assert np.array_equal(ops, r["ops"])
        assert cells > 0


def test_extend_matches_golden(oracle):
    z, rows = load_npz_rows("extend_candidates.npz")
    ref = z["ref"].tobytes()
    bases, off = z["bases"], z["offsets"]
    n_ok = 0
    for i, r in enumerate(rows):
        rd = synth.orient(bases[off[i]:off[i + 1]].tobytes(), int(r["strand"]))
        a = oracle.extend(ref, rd, int(r["loc1"]), int(r["loc2"]))
        assert a["ok"] == int(r["ok"])
        if a["ok"]:
            n_ok += 1
            assert (a["qb"], a["qe"], a["sb"], a["se"]) == (int(r["qb"]), int(r["qe"]), int(r["sb"]), int(r["se"]))
            assert a["qaln"] == r["qaln"].tobytes() and a["taln"] == r["taln"].tobytes()
    assert n_ok >= 12


def test_block_edge_cases(oracle):
    # M or N == 0 -> no alignment (xdrop_gapalign.cpp:31)
    assert oracle.block(np.zeros(1, np.uint8), 0, np.zeros(4, np.uint8), 4)[:3] == (0, 0, 0)
    # N <= 30: row 0 reaches column N (the only way the band can ever touch it)
    A = np.array([0, 1, 2, 3] * 5, dtype=np.uint8)
    s, ae, be, ops, _ = oracle.block(A, 20, A, 20)
    assert (s, ae, be) == (20, 20, 20) and len(ops) == 20 and set(ops) == {3}
    # N > 30: column N is unreachable (band growth stops at b_size < N, :147/:160)
    A = np.array([0, 1, 2, 3] * 20, dtype=np.uint8)
    s, ae, be, ops, _ = oracle.block(A, 80, A, 80)
    assert (ae, be) == (79, 79)


def test_oracle_vs_reference_random_blocks(oracle, reflib):
    rng = np.random.default_rng(99)
    for _ in range(200):
        A, B = random_block(rng)
        fwd = bool(rng.integers(0, 2))
        if not fwd:
            A, B = A[::-1].copy(), B[::-1].copy()
        o = oracle.block(A, len(A), B, len(B), fwd)
        r = reflib.block(A, len(A), B, len(B), fwd)
        assert o[:3] == r[:3] and np.array_equal(o[3], r[3])


def test_oracle_vs_reference_extend(oracle, reflib):
    d = synth.make_batch_torch(4242, 150_000, 10, 4000)
    ref = d["ref"].numpy().tobytes()
    bases, off = d["bases"].numpy(), d["offsets"].numpy()
    for i in range(10):
        rd = synth.orient(bases[off[i]:off[i + 1]].tobytes(), int(d["strand"][i]))
        a = oracle.extend(ref, rd, int(d["loc1"][i]), int(d["loc2"][i]))
        b = reflib.extend(ref, rd, int(d["loc1"][i]), int(d["loc2"][i]))
        assert a["ok"] == b["ok"] == 1
        for k in ("qb", "qe", "sb", "se", "qaln", "taln"):
            assert a[k] == b[k], k
