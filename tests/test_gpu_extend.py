"""GPU parity tests proper: the CUDA path, called through the C ABI (include/ag2_b200.h), against the
oracle on the same seeded inputs, against the committed golden vectors, and -- at sizes the oracle
cannot finish -- through size-independent properties."""
import numpy as np
import pytest

from aligngraph2_b200 import synth
from conftest import load_npz_rows

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from aligngraph2_b200.mecat2ref import Mecat2RefDevice
    d = Mecat2RefDevice(0)
    yield d
    d.close()


@pytest.fixture(autouse=True)
def pair_path(monkeypatch):
    """These tests are about the pair / lane / wide kernels: keep small candidate batches away from the low-latency
    row-parallel path the library picks for them (covered by test_small_batches_take_the_row_parallel_path)."""
    monkeypatch.setenv("AG2_NO_SMALL_BATCH", "1")


def _strings(rec, qa, sa):
    o, n = int(rec["aln_off"]), int(rec["aln_len"])
    return qa[o:o + n].tobytes(), sa[o:o + n].tobytes()


def _check_against_oracle(dev, oracle, ref, bases, off, cand):
    dev.load_reference(ref)
    dev.load_reads(bases=bases, offsets=off)
    rec, qa, sa = dev.extend(cand)
    st = dev.stats()
    cells = 0
    aligned = 0
    for i, c in enumerate(cand):
        r = int(c["read"])
        rd = synth.orient(bases[off[r]:off[r + 1]].tobytes(), int(c["strand"]))
        a = oracle.extend(ref.tobytes(), rd, int(c["loc1"]), int(c["loc2"]))
        cells += a["cells"]
        assert int(rec[i]["ok"]) == a["ok"], i
        assert (int(rec[i]["read"]), int(rec[i]["strand"]), int(rec[i]["vscore"])) == (r, int(c["strand"]), int(c["score"]))
        if a["ok"]:
            aligned += a["qe"] - a["qb"]
            got = (int(rec[i]["qb"]), int(rec[i]["qe"]), int(rec[i]["sb"]), int(rec[i]["se"]), int(rec[i]["qs"]))
            assert got == (a["qb"], a["qe"], a["sb"], a["se"], len(rd)), i
            q, s = _strings(rec[i], qa, sa)
            assert q == a["qaln"] and s == a["taln"], i
    assert st["cells"] == cells          # the kernel's own cell count is the oracle's C
    assert st["aligned"] == aligned
    return st


def test_golden_candidates(dev, oracle):
    z, rows = load_npz_rows("extend_candidates.npz")
    n = len(rows)
    cand = dev.make_candidates(np.arange(n), [int(r["strand"]) for r in rows], [int(r["loc1"]) for r in rows],
                               [int(r["loc2"]) for r in rows], score=np.arange(n) + 5)
    dev.load_reference(z["ref"])
    dev.load_reads(bases=z["bases"], offsets=z["offsets"])
    rec, qa, sa = dev.extend(cand)
    for i, r in enumerate(rows):
        assert int(rec[i]["ok"]) == int(r["ok"])
        if int(r["ok"]):
            assert (int(rec[i]["qb"]), int(rec[i]["qe"]), int(rec[i]["sb"]), int(rec[i]["se"])) == \
                   (int(r["qb"]), int(r["qe"]), int(r["sb"]), int(r["se"]))
            q, s = _strings(rec[i], qa, sa)
            assert q == r["qaln"].tobytes() and s == r["taln"].tobytes()


def test_clr_reads_match_oracle(dev, oracle):
    d = synth.make_batch_torch(20261017, 1_000_000, 96, 10000)
    cand = dev.make_candidates(np.arange(96), d["strand"].numpy(), d["loc1"].numpy(), d["loc2"].numpy(), score=9)
    st = _check_against_oracle(dev, oracle, d["ref"].numpy(), d["bases"].numpy(), d["offsets"].numpy(), cand)
    assert st["blocks"] > 96 * 15


def test_ragged_and_edge_candidates(dev, oracle):
    # read lengths from 40 to 4000 bases, seeds at position 0 and at the very end, soft-masked/N bases,
    # reads hanging over both reference ends
    rng = np.random.default_rng(7)
    ref = synth.make_reference(rng, 30_000)
    reads, cands = [], []
    for i, tl in enumerate([40, 200, 999, 1000, 1200, 2500, 4000, 1500, 1500, 1500]):
        start = {7: 0, 8: len(ref) - tl}.get(i, int(rng.integers(0, len(ref) - tl)))
        rd, _, _ = synth.make_read(rng, ref[start:start + tl + 1] if i != 8 else ref[start:], min(tl, len(ref) - start - 1) if i != 8 else tl - 1, False)
        rd = rd.copy()
        if i == 5:
            idx = rng.integers(0, len(rd), size=40)
            rd[idx[:30]] |= 0x20
            rd[idx[30:]] = ord("N")
        strand = i & 1
        given = rd.tobytes() if not strand else synth.orient(rd.tobytes(), 1)
        reads.append(np.frombuffer(given, dtype=np.uint8))
        for loc2 in (0, len(rd) // 2, len(rd) - 1, len(rd)):
            loc1 = min(len(ref), max(1, start + loc2 + 1))
            cands.append((i, strand, loc1, loc2))
    off = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    c = np.array(cands)
    cand = dev.make_candidates(c[:, 0], c[:, 1], c[:, 2], c[:, 3], score=1)
    _check_against_oracle(dev, oracle, ref, np.concatenate(reads), off, cand)


def _repeat_case(dev, seed=3):
    rng = np.random.default_rng(seed)
    unit = synth.make_reference(rng, 7)
    ref = np.concatenate([synth.make_reference(rng, 5000), np.tile(unit, 600), synth.make_reference(rng, 5000)])
    reads, cands = [], []
    for i in range(12):
        start = int(rng.integers(3000, 6000))
        rd, _, _ = synth.make_read(rng, ref[start:start + 3001], 3000, False)
        reads.append(rd)
        cands.append((i, 0, start + 1501, 1500 + int(rng.integers(-40, 40))))
    off = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    c = np.array(cands)
    return ref, np.concatenate(reads), off, dev.make_candidates(c[:, 0], c[:, 1], c[:, 2], c[:, 3])


def test_repeats_and_wide_bands(dev, oracle):
    # tandem repeats make the band outgrow the 128-column fast window: the wide kernel must take over
    ref, bases, off, cand = _repeat_case(dev)
    st = _check_against_oracle(dev, oracle, ref, bases, off, cand)
    assert st["wide_chains"] > 0 or st["interior"] > 0


def test_late_hand_overs_are_continued_at_their_block(dev, oracle):
    """A tandem repeat deep inside the reads: the pair kernel runs ~8 blocks of a direction, then the band leaves its window
    and the direction is handed over -- the consumer continues it from that block in the pair format (run_chain_resumed),
    and records, strings and the cell count must still be the reference's."""
    rng = np.random.default_rng(11)
    unit = synth.make_reference(rng, 7)
    ref = np.concatenate([synth.make_reference(rng, 9000), np.tile(unit, 500), synth.make_reference(rng, 9000)])
    reads, cands = [], []
    for i in range(16):
        start = int(rng.integers(2000, 4000))
        rd, _, _ = synth.make_read(rng, ref[start:start + 9001], 9000, False)
        strand = i & 1
        reads.append(np.frombuffer(rd.tobytes() if not strand else synth.orient(rd.tobytes(), 1), dtype=np.uint8))
        cands.append((i, strand, start + 501, 500))
    off = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    c = np.array(cands)
    cand = dev.make_candidates(c[:, 0], c[:, 1], c[:, 2], c[:, 3])
    st = _check_against_oracle(dev, oracle, ref, np.concatenate(reads), off, cand)
    assert st["lane_chains"] > 0        # directions did leave the pair kernel


def test_small_batches_take_the_row_parallel_path(dev, oracle, monkeypatch):
    """A handful of candidates (rescue extensions, second pass) runs one warp per direction with a 128-column window, and
    what leaves that window on the full-width form: same records, same strings, same cell count."""
    monkeypatch.delenv("AG2_NO_SMALL_BATCH")
    d = synth.make_batch_torch(4242, 400_000, 24, 6000)
    ref, bases, off = d["ref"].numpy(), d["bases"].numpy(), d["offsets"].numpy()
    cand = dev.make_candidates(np.arange(24), d["strand"].numpy(), d["loc1"].numpy(), d["loc2"].numpy())
    st = _check_against_oracle(dev, oracle, ref, bases, off, cand)
    assert st["lane_chains"] == 0 and st["wide_chains"] >= 48      # nothing went through the pair kernel
    # tandem repeats: bands wider than 128 columns fall back to the 736-column form
    ref, bases, off, cand = _repeat_case(dev, seed=4)
    _check_against_oracle(dev, oracle, ref, bases, off, cand)


def test_invalid_candidates_are_not_ok(dev):
    d = synth.make_batch_torch(5, 50_000, 4, 1500)
    dev.load_reference(d["ref"].numpy())
    dev.load_reads(bases=d["bases"].numpy(), offsets=d["offsets"].numpy())
    cand = dev.make_candidates([0, 9, 1, 2], [0, 0, 0, 1], [100, 100, 10**9, 50], [10, 10, 10, 10**6])
    rec, _, _ = dev.extend(cand)
    assert list(rec["ok"][1:]) == [0, 0, 0]


def test_properties_at_scale(dev):
    # 20k CLR reads (200 Mbp): too many for the oracle; check what must hold for every record
    import torch
    d = synth.make_batch_torch(77, 5_000_000, 20000, 10000, device="cuda")
    ref = d["ref"].cpu().numpy()
    bases, off = d["bases"].cpu().numpy(), d["offsets"].cpu().numpy()
    n = 20000
    cand = dev.make_candidates(np.arange(n), d["strand"].cpu().numpy(), d["loc1"].cpu().numpy(), d["loc2"].cpu().numpy())
    dev.load_reference(ref)
    dev.load_reads(bases=bases, offsets=off)
    rec, qa, sa = dev.extend(cand)
    assert rec["ok"].mean() > 0.99
    ok = rec[rec["ok"] == 1]
    # columns: every column consumes a base on at least one side, counts add up
    gap = ord("-")
    starts = ok["aln_off"]
    qnon = np.add.reduceat((qa != gap).astype(np.int64), starts)
    snon = np.add.reduceat((sa != gap).astype(np.int64), starts)
    assert np.array_equal(qnon, ok["qe"] - ok["qb"])
    assert np.array_equal(snon, ok["se"] - ok["sb"])
    assert not np.any((qa == gap) & (sa == gap))
    # strings spell the read / reference they claim to align (spot check 200 records)
    rng = np.random.default_rng(0)
    for i in rng.choice(len(ok), size=200, replace=False):
        r = ok[i]
        q, s = _strings(r, qa, sa)
        rd = synth.orient(bases[off[r["read"]]:off[r["read"] + 1]].tobytes(), int(r["strand"]))
        assert q.replace(b"-", b"") == rd[r["qb"]:r["qe"]]
        assert s.replace(b"-", b"") == ref[r["sb"]:r["se"]].tobytes()
    # the true template interval is recovered (reads were cut from the reference)
    st = d["start"].cpu().numpy()[ok["read"]]
    assert np.mean(np.abs(ok["sb"] - st) < 50) > 0.98
    # idempotence: a second run over the same resident inputs gives identical bytes
    rec2, qa2, sa2 = dev.extend(cand)
    assert np.array_equal(rec, rec2) and np.array_equal(qa, qa2) and np.array_equal(sa, sa2)


@pytest.mark.parametrize("path", ["streamed", "streamed-hostwait", "chunked"])
@pytest.mark.parametrize("order", ["read", "shuffled"])
def test_async_read_load_and_streamed_chunks(dev, monkeypatch, order, path):
    """ag2_reads_load_async + ag2_xdrop_extend_batch with host buffers: the reads go up in many pieces while ONE launch
    already extends the candidates (every direction waits for the piece that holds its read inside the kernel), and the
    results come home in many output chunks as the kernel raises their flags (test knobs make pieces and chunks small).
    Records and strings must equal the synchronous, resident run -- whatever the candidate order."""
    from aligngraph2_b200.lib import RECORD_DTYPE
    from aligngraph2_b200.mecat2ref import Mecat2RefDevice
    d = synth.make_batch_torch(99, 400_000, 600, 4000)
    ref, bases, off = d["ref"].numpy(), d["bases"].numpy(), d["offsets"].numpy()
    idx = np.arange(600)
    if order == "shuffled":
        np.random.default_rng(1).shuffle(idx)
    cand = dev.make_candidates(idx, d["strand"].numpy()[idx], d["loc1"].numpy()[idx], d["loc2"].numpy()[idx], score=7)
    dev.load_reference(ref)
    dev.load_reads(bases=bases, offsets=off)
    rec, qa, sa = dev.extend(cand)
    assert rec["ok"].mean() > 0.98

    monkeypatch.setenv("AG2_E2E_PATH", path.split("-")[0])  # both forms of the host-buffer run (ag2_xdrop_extend_batch)
    if path.endswith("hostwait"):                           # streamed, but the launch waits for the last piece of the reads
        monkeypatch.setenv("AG2_STREAM_WAIT_HOST", "1")
    monkeypatch.setenv("AG2_WS_STREAMED", str(400_000))    # ~ 30 reads per chunk
    monkeypatch.setenv("AG2_PIECE_BYTES", str(100_000))    # ~ 25 reads per piece
    dev2 = Mecat2RefDevice(0)
    try:
        dev2.load_reference(ref)
        for _ in range(2):                                  # a second batch on the same context reuses pieces and events
            rec2 = np.zeros(600, RECORD_DTYPE)
            q2, s2 = np.zeros(qa.size + 64, np.uint8), np.zeros(sa.size + 64, np.uint8)
            dev2.load_reads_async(bases, off)
            used = dev2.extend_batch_into(cand, rec2, q2, s2)
            assert used == qa.size
            assert np.array_equal(rec2, rec) and np.array_equal(q2[:used], qa) and np.array_equal(s2[:used], sa)
        st = dev2.stats()
        assert st["cells"] == dev.stats()["cells"] and st["launches"] >= 8 + 3 * 15   # really went through many output chunks
        # every other entry point waits for an asynchronous load by itself
        dev2.load_reads_async(bases, off)
        dev2.build_index(200, 0.5, 2.0)
        dev2.load_reads_async(bases, off)
        dev2.wait_reads()
    finally:
        dev2.close()


@pytest.mark.parametrize("path", ["streamed", "chunked"])
def test_packed_ops_expand_to_the_same_strings(dev, monkeypatch, path):
    """ag2_xdrop_extend_batch_packed returns the alignments as 2-bit ops (an eighth of the bytes of the two ASCII strings);
    ag2_expand_alignments rebuilds the strings on the host from the ops, the reads and the reference.  Both forms of the
    host-buffer run, many output chunks, reads of both strands with soft-masked and N bases: the rebuilt strings must be the
    strings of the plain call, byte for byte; so must the ops fetched after a resident run."""
    from aligngraph2_b200.lib import RECORD_DTYPE
    d = synth.make_batch_torch(77, 300_000, 300, 3000)
    ref, bases, off = d["ref"].numpy(), d["bases"].numpy().copy(), d["offsets"].numpy()
    rng = np.random.default_rng(2)
    for k in rng.integers(0, bases.size, size=400):     # lower case and N: the reverse strand must not complement them
        bases[k] = ord("N") if rng.random() < 0.3 else bases[k] | 0x20
    cand = dev.make_candidates(np.arange(300), d["strand"].numpy(), d["loc1"].numpy(), d["loc2"].numpy(), score=3)
    dev.load_reference(ref)
    dev.load_reads(bases=bases, offsets=off)
    rec, qa, sa = dev.extend(cand)
    assert rec["ok"].mean() > 0.9
    # resident run -> packed fetch
    rec_p, ops_p, cols = dev.fetch_packed(300)
    assert cols == qa.size and np.array_equal(rec_p, rec)
    q2, s2 = dev.expand_alignments(rec_p, ops_p, bases, off, ref, cols)
    assert np.array_equal(q2[:cols], qa) and np.array_equal(s2[:cols], sa)
    # host-buffer run, packed
    monkeypatch.setenv("AG2_E2E_PATH", path)
    monkeypatch.setenv("AG2_WS_STREAMED", str(300_000))
    rec3 = np.zeros(300, RECORD_DTYPE)
    ops3 = np.zeros(qa.size // 16 + 4096, np.uint32)
    dev.load_reads_async(bases, off)
    used = dev.extend_batch_packed_into(cand, rec3, ops3)
    assert used >= qa.size and used < qa.size + 16 * 300
    same = [k for k in RECORD_DTYPE.names if k != "aln_off"]
    assert all(np.array_equal(rec3[k], rec[k]) for k in same)       # chunks start on word boundaries: only aln_off may differ
    q3, s3 = dev.expand_alignments(rec3, ops3, bases, off, ref, used, threads=3)
    for i in np.nonzero(rec["ok"])[0]:
        o, o3, m = int(rec[i]["aln_off"]), int(rec3[i]["aln_off"]), int(rec[i]["aln_len"])
        assert q3[o3:o3 + m].tobytes() == qa[o:o + m].tobytes() and s3[o3:o3 + m].tobytes() == sa[o:o + m].tobytes(), i
