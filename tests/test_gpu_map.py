"""GPU parity of the whole per-read path (ag2_map_reads = reference_mapping()'s loop body):
the `.r` thread file written from the device results must be byte-identical to the reference binary's
(committed golden copy) and to the oracle's on other inputs."""
import lzma
import os

import numpy as np
import pytest

from aligngraph2_b200 import synth
from conftest import GOLDEN, golden_ref_outputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from aligngraph2_b200.mecat2ref import Mecat2RefDevice
    d = Mecat2RefDevice(0)
    yield d
    d.close()


def _thread_file(dev, genome: bytes, bases: bytes, offs, path, cbl=200, alpha=0.5, beta=2.0, maxc=10, num_output=1):
    from oracle.binding import MapperOracle
    dev.load_reference(np.frombuffer(MapperOracle.upper_ref(genome), dtype=np.uint8))
    dev.load_reads(bases=np.frombuffer(bases, dtype=np.uint8), offsets=offs)
    dev.build_index(cbl, alpha, beta)
    rec, qa, sa = dev.map_reads(maxc, num_output)
    dev.write_thread_file(path, rec, qa, sa, np.arange(1, len(offs)))
    return rec


def test_stress_fixture_matches_reference_thread_file(dev, tmp_path):
    z = np.load(os.path.join(GOLDEN, "mapper_stress.npz"))
    out = str(tmp_path / "1.r")
    rec = _thread_file(dev, z["genome"].tobytes(), z["bases"].tobytes(), z["offsets"].astype(np.int64), out)
    golden = golden_ref_outputs()["wrk/1.r"]
    assert open(out, "rb").read() == golden
    assert len(rec) == golden.count(b"\n") // 3


def test_clr_reads_match_oracle_thread_file(dev, tmp_path):
    from oracle.binding import MapperOracle
    d = synth.make_batch_torch(31337, 2_000_000, 400, 10000)
    genome, bases, offs = d["ref"].numpy().tobytes(), d["bases"].numpy().tobytes(), d["offsets"].numpy()
    out = str(tmp_path / "gpu.r")
    _thread_file(dev, genome, bases, offs, out)
    exp = str(tmp_path / "oracle.r")
    n, st = MapperOracle().map_batch(genome, bases, offs, np.arange(1, len(offs)), exp)
    assert n >= 399
    assert open(out, "rb").read() == open(exp, "rb").read()


def test_more_outputs_and_fewer_candidates(dev, tmp_path):
    # -b 3 -n 5 -z 10000 (AlignGraph2.py really passes -z = its -a default)
    from oracle.binding import MapperOracle
    z = np.load(os.path.join(GOLDEN, "mapper_stress.npz"))
    genome, bases, offs = z["genome"].tobytes(), z["bases"].tobytes(), z["offsets"].astype(np.int64)
    out = str(tmp_path / "gpu.r")
    _thread_file(dev, genome, bases, offs, out, cbl=10000, maxc=5, num_output=3)
    exp = str(tmp_path / "oracle.r")
    MapperOracle().map_batch(genome, bases, offs, np.arange(1, len(offs)), exp, cbl=10000, maxc=5, num_output=3)
    assert open(out, "rb").read() == open(exp, "rb").read()


def test_stress_fixture_through_every_seeding_tier(dev, tmp_path, monkeypatch):
    """The CTA-per-read seeding / rescue planning hands reads that do not fit its shared memory to a second launch with a
    larger table and, beyond that, to the one-thread-per-read path: tiny limits force reads through all three, and the
    thread file must not change."""
    z = np.load(os.path.join(GOLDEN, "mapper_stress.npz"))
    golden = golden_ref_outputs()["wrk/1.r"]
    seen = []
    for cap, cap2 in ((64, 128), (256, 1024), (64, 13824)):
        monkeypatch.setenv("AG2_SEED_CAP", str(cap))
        monkeypatch.setenv("AG2_SEED_CAP2", str(cap2))
        out = str(tmp_path / f"{cap}.r")
        _thread_file(dev, z["genome"].tobytes(), z["bases"].tobytes(), z["offsets"].astype(np.int64), out)
        assert open(out, "rb").read() == golden, (cap, cap2)
        st = dev.map_stats()
        seen.append((st["seed_overflow1"], st["seed_overflow2"]))
    assert seen[0][1] > 0 and seen[1][0] > 0, seen      # the thread path and the second launch were really used


def test_250mb_reference_matches_reference_binary(dev, tmp_path):
    """BASELINE configs[2]'s reference size against the UNMODIFIED reference binary: 320 reads (CLR templates of both
    strands, chimeras, unrelated reads, short reads) vs a seeded 250 Mb reference; `mecat2ref -t 1`'s thread file is
    committed as digests (tests/golden/map250_ref.json, gen_map250_golden.py), inputs are regenerated from the seeds."""
    import hashlib
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("gen_map250_golden", os.path.join(GOLDEN, "gen_map250_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    gold = json.load(open(gen.GOLDEN_JSON))
    ref, reads = gen.inputs(gold["seed"], gold["ref_len"], gold["n_reads"])
    assert hashlib.sha256(b"\n".join(reads)).hexdigest() == gold["reads_sha256"]      # the generator still makes the golden's inputs
    assert hashlib.sha256(ref.tobytes()).hexdigest() == gold["ref_sha256"]
    offs = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in reads], out=offs[1:])
    out = str(tmp_path / "gpu.r")
    _thread_file(dev, ref.tobytes(), b"".join(reads), offs, out)
    tf = open(out, "rb").read()
    got = gen.record_digest(tf)
    assert len(got) == len(gold["records"])
    for a, b in zip(got, gold["records"]):
        assert a == b
    assert hashlib.sha256(tf).hexdigest() == gold["thread_file_sha256"]
    st = dev.map_stats()
    assert st["n_pass2_reads"] > 0      # the golden exercises the second pass at this size (and the rescue searches: 40 clipped reads)


def test_250mb_reference_properties(dev):
    """BASELINE configs[2] scale on the reference side: a 250 Mb reference (13-mer buckets of 3.7 positions, 1.25 M vote
    blocks, positions beyond 2^27), whole per-read path.  No oracle at this size: the reads were cut from the reference, so
    every read must come back on its own strand at its own template start, and the strings must spell the read and the
    reference interval they claim."""
    import torch
    n, R = 1500, 250_000_000
    d = synth.make_batch_torch(777, R, n, 10000, device="cuda")
    ref = d["ref"].cpu().numpy()
    bases, off = d["bases"].cpu().numpy(), d["offsets"].cpu().numpy()
    start = d["start"].cpu().numpy()
    del d
    torch.cuda.empty_cache()
    dev.load_reference(ref)
    dev.load_reads(bases=bases, offsets=off)
    dev.build_index(200, 0.5, 2.0)
    rec, qa, sa = dev.map_reads(10, 1)
    assert len(np.unique(rec["read"])) >= 0.98 * n
    assert np.all(rec["se"] <= R) and np.all(rec["sb"] >= 0) and np.all(rec["se"] > rec["sb"])
    # the longest record of every read: right strand, right place, (nearly) the whole read
    order = np.lexsort((-(rec["qe"] - rec["qb"]), rec["read"]))
    best = rec[order][np.unique(rec["read"][order], return_index=True)[1]]
    assert np.mean(best["strand"] == (best["read"] & 1)) > 0.98
    assert np.mean(np.abs(best["sb"] - start[best["read"]]) < 50) > 0.95
    assert np.mean((best["qe"] - best["qb"]) > 0.9 * best["qs"]) > 0.95
    rng = np.random.default_rng(5)
    for i in rng.choice(len(best), size=60, replace=False):
        r = best[i]
        o, m = int(r["aln_off"]), int(r["aln_len"])
        q, s = qa[o:o + m].tobytes(), sa[o:o + m].tobytes()
        rd = synth.orient(bases[off[r["read"]]:off[r["read"] + 1]].tobytes(), int(r["strand"]))
        assert len(q) == len(s) and q.replace(b"-", b"") == rd[r["qb"]:r["qe"]]
        assert s.replace(b"-", b"") == ref[r["sb"]:r["se"]].tobytes()
