"""PAGraph traversal and assembly of the walks: ag2_pg_travel -- a host stage by design -- on the golden
graph of the UNMODIFIED reference classes must write, byte for byte, the files PAssembly::testTravel5 writes
(tests/golden/pagraph_travel.tar.xz: t1/, t8/ from oracle/_ref/pagraph_dump; t1/ == what `pagraph -t 1` writes), and the
same on fresh data where the reference binaries exist.  No GPU needed: the graph comes from the fixture."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
sys.path.insert(0, os.path.dirname(__file__))
import gen_pagraph_golden as gen_small  # noqa: E402
import gen_pagraph_travel_golden as gen  # noqa: E402
import pg_fixture as fx  # noqa: E402


@pytest.fixture(scope="module")
def lib():
    from aligngraph2_b200 import build
    build.build()


def _travel_all(d, out, threads, eps=10, min_len=50):
    from aligngraph2_b200 import pagraph
    codes, k, ctgs, refs, blocks = fx.load_inputs(d)
    graphs = fx.graphs_from_dump(os.path.join(d, "graph.txt"), len(codes))
    os.makedirs(out, exist_ok=True)
    ok = []
    for b, use in enumerate(blocks):
        ok.append(pagraph.travel(graphs[b], codes, k, ctgs, refs, use, pagraph.travel_params(eps, min_len, threads), out, f"{b}_"))
    return ok, ctgs[0]


@pytest.fixture(scope="module")
def chain(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("pagraph_travel"))
    gen.unpack(d)
    return d


@pytest.mark.parametrize("threads", [1, 8])
def test_walk_files_equal_reference(lib, chain, threads):
    out = os.path.join(chain, f"mine{threads}")
    ok, names = _travel_all(chain, out, threads)
    want = os.path.join(chain, f"t{threads}")
    assert fx.compare_dirs(out, want, skip=("contig.txt",)) == []
    assert sorted({names[i] for b in ok for i in b}) == sorted(open(os.path.join(want, "contig.txt")).read().split())
    # the fixture reaches what it was built for: a chain of three contigs written as one FASTA
    con = open(os.path.join(want, "0_0_0.con")).read().split("\n")
    assert [l.split("\t")[:2] for l in con[1:4]] == [["ctg0", "FORWARD"], ["ctg1", "FORWARD"], ["ctg2", "REV"]]


def test_start_vertex_count_changes_the_walks(chain):
    """min(t, 8) start vertices per round (PAlgorithm.cpp:146) is visible in the golden files, so both cases are pinned."""
    assert fx.compare_dirs(os.path.join(chain, "t1"), os.path.join(chain, "t8")) != []


def test_small_fixture_against_reference_binary(lib, tmp_path):
    """The two-reference fixture of the graph build (reverse-used contig, unaligned tail): needs the reference binary."""
    ref = os.path.join(gen_small.REFDIR, "pagraph_dump")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/pagraph_dump not built on this box")
    d = str(tmp_path)
    gen_small.unpack(d)
    for threads, min_len in ((1, 50), (4, 3000)):
        gen.run_travel(d, threads, f"ref{threads}", min_len=min_len)
        _travel_all(d, os.path.join(d, f"mine{threads}"), threads, min_len=min_len)
        assert fx.compare_dirs(os.path.join(d, f"mine{threads}"), os.path.join(d, f"ref{threads}"), skip=("contig.txt",)) == []
        assert len(os.listdir(os.path.join(d, f"ref{threads}"))) >= 4


def test_rejects_bad_arguments(lib, chain):
    from aligngraph2_b200 import pagraph
    from aligngraph2_b200.lib import Ag2Error
    codes, k, ctgs, refs, blocks = fx.load_inputs(chain)
    g = fx.graphs_from_dump(os.path.join(chain, "graph.txt"), len(codes))[0]
    with pytest.raises(Ag2Error):                                    # contig index outside the database
        pagraph.travel(g, codes, k, ctgs, refs, [(7, True)], pagraph.travel_params(), chain, "x_")
    with pytest.raises(Ag2Error):                                    # output directory does not exist
        pagraph.travel(g, codes, k, ctgs, refs, blocks[0], pagraph.travel_params(), os.path.join(chain, "nowhere"), "x_")


def test_pagraph_executable_without_gpu_exits_1(lib, chain):
    """The drop-in `pagraph` needs the device for the graph build: no GPU -> exit 1, nothing is walked on a CPU-built graph."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from aligngraph2_b200 import build
    exe = build.build_host(name="pagraph")
    r = subprocess.run([exe, "-t", "1", "-r", "dummy", "-k", "solid.bin", "-c", "ctg.fasta", "-R", "ref.fasta", "-p", ".", "-a", "c2r.ref",
                        "-o", ".", "-r", "50", "--epsilon", "10", "-v", "2"], cwd=chain, capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU path" in r.stderr
    assert subprocess.run([exe], capture_output=True).returncode == 0
    assert subprocess.run([exe, "--bogus", "1"], capture_output=True).returncode == 1
