"""SURVEY row N2 ("next"), first step only: the C restatement of vanilla MECAT2's DiffAligner (oracle/ag2_diff.c) is pinned
against the unmodified reference sources (oracle/_ref/libref_mecat_vanilla.so) -- on the committed golden vectors and,
where the reference library is present, on fresh random inputs.  No product code implements this row yet."""
import numpy as np
import pytest

from conftest import load_npz_rows


@pytest.fixture(scope="module")
def diff_oracle():
    from oracle import binding
    binding.build(ref=False)
    return binding.DiffOracle()


@pytest.fixture(scope="module")
def diff_ref():
    import os
    from oracle import binding
    if not binding.have_diff_ref():
        if os.path.isdir("/root/reference/thirdparty/mecat"):
            binding.build(ref=True)
        else:
            pytest.skip("oracle/_ref/libref_mecat_vanilla.so not built (no /root/reference on this box)")
    r = binding.DiffRef()
    yield r
    r.close()


def test_golden_blocks(diff_oracle):
    _, rows = load_npz_rows("diff_blocks.npz")
    assert len(rows) >= 40
    for r in rows:
        a = diff_oracle.block(r["q"], r["t"], int(r["fwd"]))
        got = [a["rc"], a["q_s"], a["q_e"], a["t_s"], a["t_e"], a["dist"], a["n"]]
        assert got == [int(v) for v in r["res"]]
        assert a["qstr"] == r["qstr"].tobytes() and a["tstr"] == r["tstr"].tobytes()


def test_golden_extensions(diff_oracle):
    _, rows = load_npz_rows("diff_go.npz")
    reached_end = 0
    for r in rows:
        qs, ts, ok, qoff, qend, toff, tend, n = (int(v) for v in r["res"])
        a = diff_oracle.go(r["q"], qs, r["t"], ts, 0)
        assert (a["ok"], a["qoff"], a["qend"], a["toff"], a["tend"], a["aln_size"]) == (ok, qoff, qend, toff, tend, n)
        assert a["qaln"] == r["qaln"].tobytes() and a["taln"] == r["taln"].tobytes()
        # what must hold for any alignment: the strings spell the sequences they cover
        letters = np.frombuffer(b"ACGT", dtype=np.uint8)
        assert a["qaln"].replace(b"-", b"") == letters[r["q"][qoff:qend]].tobytes()
        assert a["taln"].replace(b"-", b"") == letters[r["t"][toff:tend]].tobytes()
        reached_end += qend == len(r["q"])
    assert reached_end >= 5


def _mutate(t, rng):
    out = []
    for b in t:
        u = rng.random()
        if u < 0.09:
            out.append(int(rng.integers(0, 4)))
            out.append(int(b))
        elif u < 0.13:
            continue
        elif u < 0.15:
            out.append(int((b + 1 + rng.integers(0, 3)) % 4))
        else:
            out.append(int(b))
    return np.array(out, dtype=np.uint8)


def test_fresh_blocks_against_the_reference(diff_oracle, diff_ref):
    rng = np.random.default_rng(99)
    for it in range(150):
        n = int(rng.integers(1, 720))
        t = rng.integers(0, 4, n).astype(np.uint8)
        if it % 9 == 0:
            t = np.tile(rng.integers(0, 4, 6).astype(np.uint8), n)[:n]
        q = rng.integers(0, 4, int(rng.integers(1, 600))).astype(np.uint8) if it % 7 == 0 else _mutate(t, rng)
        if len(q) == 0:
            continue
        fwd = it & 1
        assert diff_oracle.block(q, t, fwd) == diff_ref.block(q, t, fwd), it


def test_fresh_extensions_against_the_reference(diff_oracle, diff_ref):
    rng = np.random.default_rng(7)
    for it in range(16):
        n = int(rng.integers(50, 6000))
        t = rng.integers(0, 4, n).astype(np.uint8)
        q = _mutate(t, rng)
        pad = rng.integers(0, 4, int(rng.integers(0, 300))).astype(np.uint8)
        T = np.concatenate([pad, t, pad])
        ts = int(rng.integers(0, n + 1))
        qs = min(len(q), int(ts * len(q) / max(1, n)))
        assert diff_oracle.go(q, qs, T, ts + len(pad), 0) == diff_ref.go(q, qs, T, ts + len(pad), 0), it


def test_large_blocks_against_the_reference(diff_oracle):
    """DiffAligner(1): 1000-base segments (diff_gapalign.h:25-35; mecat2ref itself constructs DiffAligner(0))."""
    import os
    from oracle import binding
    if not binding.have_diff_ref():
        pytest.skip("oracle/_ref/libref_mecat_vanilla.so not built")
    ref = binding.DiffRef(1)
    try:
        rng = np.random.default_rng(5)
        for it in range(4):
            n = int(rng.integers(2500, 5000))
            t = rng.integers(0, 4, n).astype(np.uint8)
            q = _mutate(t, rng)
            ts = int(rng.integers(0, n + 1))
            qs = min(len(q), int(ts * len(q) / max(1, n)))
            assert diff_oracle.go(q, qs, t, ts, 0, large_block=1) == ref.go(q, qs, t, ts, 0), it
    finally:
        ref.close()
