"""PAGraph A-Bruijn build: the restatement oracle/ag2_pagraph.cpp against the UNMODIFIED reference
sources (oracle/_ref/pagraph_dump) -- on the committed fixture (runs anywhere) and on fresh data with other seeds and
parameters (only where the reference binaries exist)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import gen_pagraph_golden as gen  # noqa: E402


@pytest.fixture(scope="module")
def small(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("pagraph_small"))
    gen.unpack(d)
    return d


def test_oracle_reproduces_committed_reference_dump(small):
    from oracle import binding
    binding.build(ref=False)
    got = binding.pagraph_dump(small, "oracle.txt", eps=10, cov=2)
    want = open(os.path.join(small, "graph.txt"), "rb").read()
    assert got == want
    cfgs = binding.parse_graph_dump(got)
    assert len(cfgs) == 2 and min(len(c) for c in cfgs) > 10000
    # the data reaches what it was built for: merged positions (count > 1), (0, x) and (x, 0) entries, several edges
    flat = [p for c in cfgs for v in c.values() for p in v[1]]
    assert any(p[2] > 1 for p in flat) and any(p[0] == 0 for p in flat) and any(p[1] == 0 for p in flat)
    assert any(len(v[2]) > 2 for c in cfgs for v in c.values())


@pytest.mark.parametrize("eps,cov", [(0, 1), (25, 3)])
def test_oracle_matches_reference_binary_other_parameters(small, eps, cov):
    from oracle import binding
    if not os.path.exists(binding.REF_PAGRAPH_DUMP):
        pytest.skip("oracle/_ref/pagraph_dump not built on this box")
    want = gen.run_dump(small, eps=eps, cov=cov, out=f"ref_{eps}_{cov}.txt")
    got = binding.pagraph_dump(small, f"orc_{eps}_{cov}.txt", eps=eps, cov=cov)
    assert got == want


def test_oracle_matches_reference_binary_fresh_data(tmp_path):
    from oracle import binding
    if not all(os.path.exists(os.path.join(gen.REFDIR, b)) for b in ("pagraph_dump", "mecat2ref", "kmer_counter")):
        pytest.skip("oracle/_ref binaries not built on this box")
    d = str(tmp_path)
    gen.build(d, seed=5, genome_len=40000, n_reads=60, tlen=3000, k=8)
    want = gen.run_dump(d, eps=10, cov=1)
    got = binding.pagraph_dump(d, "orc.txt", eps=10, cov=1)
    assert got == want
