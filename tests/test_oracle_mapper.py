"""Pins oracle/ag2_mapper.c: its `.r` thread file must be byte-identical
to what the UNMODIFIED reference binary writes -- against the committed golden copy
(tests/golden/mapper_stress.r.xz, made by tests/golden/gen_mapper_golden.py) and, where oracle/_ref
exists, against a fresh run of the binary on another seed."""
import importlib.util
import lzma
import os
import shutil
import tempfile

import numpy as np
import pytest

from conftest import GOLDEN, golden_ref_outputs


@pytest.fixture(scope="module")
def mapper():
    from oracle import binding
    binding.build(ref=False)
    return binding.MapperOracle()


def _gen():
    spec = importlib.util.spec_from_file_location("gen_mapper_golden", os.path.join(GOLDEN, "gen_mapper_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_thread_file_matches_golden(mapper, tmp_path):
    z = np.load(os.path.join(GOLDEN, "mapper_stress.npz"))
    offs = z["offsets"]
    out = str(tmp_path / "1.r")
    n, st = mapper.map_batch(z["genome"].tobytes(), z["bases"].tobytes(), offs, np.arange(1, len(offs)), out)
    golden = golden_ref_outputs()["wrk/1.r"]
    assert open(out, "rb").read() == golden
    assert n == golden.count(b"\n") // 3
    # the fixture reaches the branches a uniform random genome never does
    assert st["insert_loc"] > 1000 and st["rescue_extensions"] > 0 and st["multi_candidate_reads"] > 10
    assert st["pass2_reads"] > 5 and st["votes_ne_1"] > 100


def test_thread_file_matches_reference_binary(mapper, tmp_path):
    from oracle import binding
    if not os.path.exists(binding.REF_BIN):
        pytest.skip("oracle/_ref/mecat2ref not built on this box")
    gen = _gen()
    chroms, reads = gen.build(seed=1234)
    reads = reads[:45]
    d = tempfile.mkdtemp(prefix="m2r_", dir=str(tmp_path))
    gen.write_inputs(d, chroms, reads)
    ref_r = gen.run_reference(d)
    offs = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum([len(x) for x in reads], out=offs[1:])
    out = os.path.join(d, "oracle.r")
    mapper.map_batch(b"".join(s for _, s in chroms), b"".join(reads), offs, np.arange(1, len(reads) + 1), out)
    assert open(out, "rb").read() == ref_r
    shutil.rmtree(d, ignore_errors=True)
