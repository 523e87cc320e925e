"""PAGraph A-Bruijn build on the GPU through the C ABI of include/ag2_pagraph.h: the graph of
every config block must equal, byte for byte in dump form, the graph the UNMODIFIED reference sources build
(tests/golden/pagraph_small.tar.xz: graph.txt from oracle/_ref/pagraph_dump) and the oracle's for other parameters."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import gen_pagraph_golden as gen  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("pagraph_small"))
    gen.unpack(d)
    return d


def _job(d):
    from aligngraph2_b200 import build, pagraph
    build.build()
    j = lambda n: os.path.join(d, n)
    return pagraph.Job(j("solid.bin"), j("ctg.fasta"), j("ref.fasta"), d, j("c2r.ref"))


def _gpu_dump(d, out, eps, cov, staged=False):
    from aligngraph2_b200 import pagraph
    job = _job(d)
    p = pagraph.default_params(eps, cov)
    stats = []
    for b in range(job.n_blocks):
        job.load_block(b)
        if staged:
            job.extract(p)
            stats.append(job.join(p).as_dict())
        else:
            stats.append(job.build(p).as_dict())
        job.dump(b, os.path.join(d, out), append=b > 0)
    job.close()
    return open(os.path.join(d, out), "rb").read(), stats


def _explain(got, want):
    g, w = got.split(b"\n"), want.split(b"\n")
    for i, (a, b) in enumerate(zip(g, w)):
        if a != b:
            return f"line {i}: got {a[:300]!r} want {b[:300]!r} ({len(g)} vs {len(w)} lines)"
    return f"{len(g)} vs {len(w)} lines"


def test_graph_equals_reference_dump(small):
    want = open(os.path.join(small, "graph.txt"), "rb").read()
    got, stats = _gpu_dump(small, "gpu.txt", 10, 2)
    assert got == want, _explain(got, want)
    assert len(stats) == 2 and all(s["tuples"][0] > 1000 and s["tuples"][1] > 1000 and s["edges"] > 1000 for s in stats)
    # "merge pos" really merged something, and the kernels ran
    assert all(s["positions"] < sum(s["tuples"]) and s["launches"] > 10 for s in stats)


@pytest.mark.parametrize("eps,cov", [(0, 1), (25, 3), (1000, 1)])
def test_graph_equals_oracle_other_parameters(small, eps, cov):
    from oracle import binding
    binding.build(ref=False)
    want = binding.pagraph_dump(small, f"orc_{eps}_{cov}.txt", eps=eps, cov=cov)
    got, _ = _gpu_dump(small, f"gpu_{eps}_{cov}.txt", eps, cov, staged=True)
    assert got == want, _explain(got, want)


def test_partitioned_streams_give_the_same_graph(small):
    """The multi-GPU path on one GPU: extract, stable partition by owner (3 owners), re-import the partitioned streams in
    owner order (what the all-to-all delivers to a single rank is a permutation that keeps every vertex's order), join."""
    from aligngraph2_b200 import pagraph
    want = open(os.path.join(small, "graph.txt"), "rb").read()
    job = _job(small)
    p = pagraph.default_params(10, 2)
    for b in range(job.n_blocks):
        job.load_block(b)
        job.extract(p)
        counts = job.partition(3)
        nt, tp, ne, ep = job.stream_pointers()
        assert counts[0].sum() == nt and counts[1].sum() == ne and (counts > 0).all()
        job.import_streams(nt, tp, ne, ep)
        job.join(p)
        job.dump(b, os.path.join(small, "part.txt"), append=b > 0)
    job.close()
    got = open(os.path.join(small, "part.txt"), "rb").read()
    assert got == want, _explain(got, want)


def test_read_sharded_ranks_cover_the_graph(small):
    """Two read shards built independently (rank 0/2 and 1/2), streams concatenated on the host in rank order and joined:
    the result of the exchange without the network."""
    import torch
    from aligngraph2_b200 import pagraph
    want = open(os.path.join(small, "graph.txt"), "rb").read()
    p = pagraph.default_params(10, 2)
    jobs = [_job(small) for _ in range(2)]
    out = os.path.join(small, "shard.txt")
    for b in range(jobs[0].n_blocks):
        parts = []
        for r, job in enumerate(jobs):
            job.load_block(b, r, 2)
            job.extract(p)
            nt, tp, ne, ep = job.stream_pointers()
            dev = torch.device("cuda", 0)
            parts.append(([pagraph._device_tensor(x, nt, dev).clone() for x in tp], [pagraph._device_tensor(x, ne, dev).clone() for x in ep]))
        tup = [torch.cat([parts[0][0][i], parts[1][0][i]]) for i in range(3)]
        edg = [torch.cat([parts[0][1][i], parts[1][1][i]]) for i in range(3)]
        torch.cuda.synchronize()
        jobs[0].import_streams(tup[0].numel(), [t.data_ptr() for t in tup], edg[0].numel(), [t.data_ptr() for t in edg])
        jobs[0].join(p)
        jobs[0].dump(b, out, append=b > 0)
    for job in jobs:
        job.close()
    got = open(out, "rb").read()
    assert got == want, _explain(got, want)


@pytest.mark.parametrize("n_handles", [2, 3])
def test_group_exchange_in_one_process(small, n_handles):
    """ag2_pg_group_exchange: the reads sharded over n handles of ONE process (one per visible GPU, round robin -- on a
    one-GPU box they share the device and the peer copies are device copies), segments pushed straight into the owner's
    buffers, every handle joins its vertex range; the merged tables must be the reference's dump."""
    import torch
    from aligngraph2_b200 import pagraph
    want = open(os.path.join(small, "graph.txt"), "rb").read()
    p = pagraph.default_params(10, 2)
    ndev = max(1, torch.cuda.device_count())
    j = lambda n: os.path.join(small, n)
    jobs = [pagraph.Job(j("solid.bin"), j("ctg.fasta"), j("ref.fasta"), small, j("c2r.ref"), device=r % ndev) for r in range(n_handles)]
    codes = jobs[0].codes()
    got = b""
    for b in range(jobs[0].n_blocks):
        stats = pagraph.build_group(jobs, b, p)
        assert sum(1 for st in stats if st.positions > 0) >= 2          # more than one owner really holds vertices
        g = pagraph.merge_graphs([job.graph() for job in jobs])
        got += pagraph.graph_dump_text(g, codes, b, jobs[0].block_ref(b))
        # the same merge on the device, into handle 0 (what the multi-GPU `pagraph` executable traverses)
        pagraph.gather_group(jobs)
        jobs[0].dump(b, os.path.join(small, f"group{n_handles}.txt"), append=b > 0)
    for job in jobs:
        job.close()
    assert got == want, _explain(got, want)
    got2 = open(os.path.join(small, f"group{n_handles}.txt"), "rb").read()
    assert got2 == want, _explain(got2, want)


def test_two_gpus_nccl_exchange(small):
    """Reads sharded over 2 GPUs, NCCL all-to-all of the vertex tuples by owner, all-gather of the merged tables:
    the dump rank 0 writes equals the reference's.  Needs a 2-GPU box (gpurun --gpus 2)."""
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from aligngraph2_b200 import build
    build.build()
    j = lambda n: os.path.join(small, n)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", "-m", "aligngraph2_b200.pagraph_dist_main", "-k", j("solid.bin"), "-c", j("ctg.fasta"),
           "-R", j("ref.fasta"), "-p", small, "-a", j("c2r.ref"), "-o", j("dist.txt"), "--epsilon", "10", "-v", "2"]
    r = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    got = open(j("dist.txt"), "rb").read()
    want = open(j("graph.txt"), "rb").read()
    assert got == want, _explain(got, want)


def test_synthetic_set_with_gpu_kmer_counter_equals_oracle(tmp_path):
    """A bigger, generated input set (aligngraph2_b200.synth_pg), solid k-mers from this package's kmer_counter kernels
    (row B1), k = 12: GPU graph == oracle graph."""
    from aligngraph2_b200 import pagraph, synth_pg
    from oracle import binding
    binding.build(ref=False)
    d = str(tmp_path)
    synth_pg.make_input_set(d, 77, 600_000, 1500, tlen=5000, n_ctg=5)
    words = synth_pg.solid_words_from_reads(d, 12, 0.2, 0)
    assert words[0] == 12 and len(words) > 1000
    want = binding.pagraph_dump(d, "orc.txt", eps=10, cov=2)
    j = lambda n: os.path.join(d, n)
    job = pagraph.Job(j("solid.bin"), j("ctg.fasta"), j("ref.fasta"), d, j("c2r.ref"))
    job.load_block(0)
    st = job.build(pagraph.default_params(10, 2))
    job.dump(0, j("gpu.txt"))
    job.close()
    got = open(j("gpu.txt"), "rb").read()
    assert got == want, _explain(got, want)
    assert st.positions > 10000 and st.edges > 10000


# ---- row B9 end to end: the drop-in `pagraph` executable (GPU build + host traversal) ---------------------------------
@pytest.fixture(scope="module")
def chain(tmp_path_factory):
    import gen_pagraph_travel_golden as gen_travel
    d = str(tmp_path_factory.mktemp("pagraph_travel"))
    gen_travel.unpack(d)
    return d


def _compare_dirs(got, want):
    bad = []
    for n in sorted(set(os.listdir(got)) | set(os.listdir(want))):
        a, b = os.path.join(got, n), os.path.join(want, n)
        if not os.path.exists(a) or not os.path.exists(b) or open(a, "rb").read() != open(b, "rb").read():
            bad.append(n)
    return bad


def test_chain_fixture_graph_equals_reference_dump(chain):
    want = open(os.path.join(chain, "graph.txt"), "rb").read()
    got, _ = _gpu_dump(chain, "gpu.txt", 10, 2)
    assert got == want, _explain(got, want)


@pytest.mark.parametrize("devices", ["0,0,0", None])
def test_drop_in_pagraph_executable_sharded_over_devices(chain, devices):
    """The same executable with the graph build sharded over several handles (AG2_DEVICES names GPU 0 three times on a
    one-GPU box; None = every visible GPU): reads cut into contiguous ranges, ag2_pg_group_exchange, per-range joins,
    ag2_pg_group_gather into the first handle, traversal there -- no file may differ from the one-GPU run's."""
    import subprocess
    from aligngraph2_b200 import build
    build.build()
    exe = build.build_host(name="pagraph")
    out = os.path.join(chain, "exe_dev" + (devices or "all").replace(",", ""))
    os.makedirs(out, exist_ok=True)
    env = {k: v for k, v in os.environ.items() if k != "AG2_DEVICES"}
    if devices:
        env["AG2_DEVICES"] = devices
    r = subprocess.run([exe, "-t", "1", "-r", "dummy", "-k", "solid.bin", "-c", "ctg.fasta", "-R", "ref.fasta", "-p", ".", "-a", "c2r.ref",
                        "-o", out, "-r", "50", "--epsilon", "10", "-v", "2"], cwd=chain, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-2000:]
    assert _compare_dirs(out, os.path.join(chain, "t1")) == []


@pytest.mark.parametrize("threads", [1, 8])
def test_drop_in_pagraph_executable(chain, threads):
    """bin/pagraph with the argv of AlignGraph2.py:414-427 (incl. the second `-r`): every file it writes -- walks, FASTA,
    .con, .help, contig.txt -- equals what the reference classes write (t1/ is also what `pagraph -t 1` itself writes)."""
    import subprocess
    from aligngraph2_b200 import build
    build.build()
    exe = build.build_host(name="pagraph")
    out = os.path.join(chain, f"exe{threads}")
    os.makedirs(out, exist_ok=True)
    r = subprocess.run([exe, "-t", str(threads), "-r", "dummy", "-k", "solid.bin", "-c", "ctg.fasta", "-R", "ref.fasta", "-p", ".", "-a", "c2r.ref",
                        "-o", out, "-r", "50", "--epsilon", "10", "-v", "2"], cwd=chain, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-2000:]
    assert _compare_dirs(out, os.path.join(chain, f"t{threads}")) == []
    assert os.path.getsize(os.path.join(out, "0_0_0.fasta")) > 20000


def test_drop_in_pagraph_against_reference_binary(small):
    """Two config blocks, two references, a contig used in reverse: against the unmodified `pagraph -t 1` run on this box."""
    import subprocess
    from aligngraph2_b200 import build
    ref_exe = os.path.join(gen.REFDIR, "pagraph")
    if not os.path.exists(ref_exe):
        pytest.skip("oracle/_ref/pagraph not built on this box")
    build.build()
    exe = build.build_host(name="pagraph")
    for name, prog in (("ref_out", ref_exe), ("gpu_out", exe)):
        os.makedirs(os.path.join(small, name), exist_ok=True)
        r = subprocess.run([prog, "-t", "1", "-r", "dummy", "-k", "solid.bin", "-c", "ctg.fasta", "-R", "ref.fasta", "-p", ".", "-a", "c2r.ref",
                            "-o", name, "-r", "50", "--epsilon", "10", "-v", "2"], cwd=small, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
    assert _compare_dirs(os.path.join(small, "gpu_out"), os.path.join(small, "ref_out")) == []
    assert len(os.listdir(os.path.join(small, "ref_out"))) == 4


def test_pagraph_executable_usage_errors(chain):
    import subprocess
    from aligngraph2_b200 import build
    build.build()
    exe = build.build_host(name="pagraph")
    assert subprocess.run([exe], capture_output=True).returncode == 0                      # no arguments: help, 0 (pagraph.cpp:88-91)
    assert subprocess.run([exe, "--bogus", "1"], capture_output=True).returncode == 1       # parse error: 1 (:98-102)
    r = subprocess.run([exe, "-k", "missing.bin", "-c", "ctg.fasta", "-R", "ref.fasta", "-p", ".", "-a", "c2r.ref", "-o", "."], cwd=chain,
                       capture_output=True, text=True)
    assert r.returncode == 1 and "missing.bin" in r.stderr
