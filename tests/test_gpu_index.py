"""GPU parity of the index build and of seeding + candidate scoring (A5-A7): the CUDA
path through the C ABI against oracle/ag2_mapper.c (itself pinned byte-for-byte to the reference binary)."""
import os

import numpy as np
import pytest

from aligngraph2_b200 import synth
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from aligngraph2_b200.mecat2ref import Mecat2RefDevice
    d = Mecat2RefDevice(0)
    yield d
    d.close()


def _check(dev, genome: bytes, bases: bytes, offs, cbl=200, alpha=0.5, beta=2.0, maxc=10, passes=(0, 1)):
    from oracle.binding import IndexOracle, MapperOracle
    io = IndexOracle(genome, bases, offs, cbl, alpha, beta, maxc)
    dev.load_reference(np.frombuffer(MapperOracle.upper_ref(genome), dtype=np.uint8))
    dev.load_reads(bases=np.frombuffer(bases, dtype=np.uint8), offsets=offs)
    dev.build_index(cbl, alpha, beta)
    ix = dev.fetch_index()
    assert np.array_equal(ix["rcnt"], io.rcnt)                     # A2: masked read 13-mer histogram
    assert np.array_equal(ix["cnt"], io.cnt)                       # A3: masked reference counts
    assert np.array_equal(ix["off"], io.off)
    assert np.array_equal(ix["pos"], io.pos)                       #     CSR, ascending inside every bucket
    assert ix["nblk"] == io.nblk
    assert np.array_equal(ix["kcount"], io.kcount)                 #     per-block read-count sums
    assert np.array_equal(ix["vote"].view(np.uint32), io.vote.view(np.uint32))   # A4: bit-identical floats
    total = 0
    for p in passes:
        cands, ncand = dev.seed_candidates(p, maxc)
        for r in range(len(offs) - 1):
            exp = io.candidates(bases[offs[r]:offs[r + 1]], p)
            got = [tuple(int(cands[r, i][k]) for k in ("loc1", "loc2", "left1", "left2", "right1", "right2", "score", "num1",
                                                       "num2", "chain")) for i in range(ncand[r])]
            assert got == exp, (p, r)
            total += len(exp)
    return ix, total


def test_stress_fixture(dev):
    z = np.load(os.path.join(GOLDEN, "mapper_stress.npz"))
    ix, total = _check(dev, z["genome"].tobytes(), z["bases"].tobytes(), z["offsets"].astype(np.int64))
    assert total > 500
    assert (ix["vote"][:ix["nblk"]] != 1.0).sum() > 100            # the alpha/beta weighting is exercised
    assert (ix["cnt"] == 0).sum() > 0


def test_clr_reads_1mb(dev):
    d = synth.make_batch_torch(99, 1_000_000, 120, 10000)
    _check(dev, d["ref"].numpy().tobytes(), d["bases"].numpy().tobytes(), d["offsets"].numpy(), cbl=200, passes=(0,))


def test_other_parameters(dev):
    # default -z (10000, what AlignGraph2.py really passes: SURVEY F3), other alpha/beta, -n 4
    z = np.load(os.path.join(GOLDEN, "mapper_stress.npz"))
    _check(dev, z["genome"].tobytes(), z["bases"].tobytes(), z["offsets"].astype(np.int64), cbl=10000, alpha=0.3, beta=1.5,
           maxc=4, passes=(0,))


def test_seeding_in_many_launches(dev, monkeypatch):
    # reads that overflow both CTA-per-read launches take the one-thread-per-read path, whose block tables share a scratch
    # arena; a small arena cuts them into many launches (a few reads each), which must not change any candidate
    monkeypatch.setenv("AG2_SEED_CAP", "64")
    monkeypatch.setenv("AG2_SEED_CAP2", "64")
    monkeypatch.setenv("AG2_SEED_SCRATCH", str(300_000))
    d = synth.make_batch_torch(123, 1_000_000, 60, 10000)
    _check(dev, d["ref"].numpy().tobytes(), d["bases"].numpy().tobytes(), d["offsets"].numpy(), cbl=200, passes=(0, 1))
