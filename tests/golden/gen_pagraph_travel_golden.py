"""Fixture for the A-Bruijn traversal and its golden output files from the UNMODIFIED reference
sources (oracle/_ref/pagraph_dump = src/tools/*.cpp + oracle/ref_pagraph_dump.cpp, which calls PAssembly::testTravel5 the
way run2 does).

  python tests/golden/gen_pagraph_travel_golden.py       # needs oracle/_ref/{mecat2ref,kmer_counter,pagraph_dump,pagraph}

Writes tests/golden/pagraph_travel.tar.xz:
  the pagraph input set (reads.fq ctg.fasta ref.fasta solid.bin r2c.ref r2r.ref c2r.ref config.txt),
  graph.txt      the dump of the graph after PositionProcessor::process() (graph built with one thread),
  t1/ and t8/    everything testTravel5 + run2 write (<block>_<ctg>_<o>.txt, .fasta, .help, .con, contig.txt) with
                 threadNum 1 and 8 (start vertices per round = min(threadNum, 8), PAlgorithm.cpp:146).
t1/ is checked to be what the unmodified `pagraph -t 1` binary writes for the same input.

The data is built to reach what pagraph_small does not: walks that leap from one contig to the next over a gap bridged by
reads and the reference (chains of two and three contigs -> inDegrees, the union-find, "connected" FASTA output), a
contig walked in reverse orientation, a contig that is extended past its end, several start vertices per round.
"""
import io
import lzma
import os
import shutil
import subprocess
import sys
import tarfile
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from aligngraph2_b200 import synth  # noqa: E402
import gen_pagraph_golden as base  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = base.REFDIR


def build(d, seed=8, genome_len=24000, n_reads=260, tlen=3000, k=10, gap=800):
    rng = np.random.default_rng(seed)
    G = synth.make_reference(rng, genome_len).tobytes()
    refs = [("chrA", base.noisy(rng, G, 0.05, 0.3, 0.3))]
    a0, a1 = genome_len // 32, genome_len * 29 // 100
    b0, b1 = a1 + gap, genome_len * 60 // 100
    c0, c1 = b1 + gap, genome_len * 90 // 100
    ctgs = [("ctg0", G[a0:a1]), ("ctg1", G[b0:b1]), ("ctg2", base.rc(G[c0:c1]))]
    reads = []
    for i in range(n_reads):
        s = int(rng.integers(0, genome_len - tlen))
        rd = base.noisy(rng, G[s:s + tlen], 0.15)
        if i % 2:
            rd = base.rc(rd)
        reads.append(rd)
    with open(os.path.join(d, "reads.fq"), "wb") as f:
        for i, rd in enumerate(reads):
            f.write(b"@%d\n" % (i + 1) + rd + b"\n+\n" + b"I" * len(rd) + b"\n")
    base.write_fasta(os.path.join(d, "ctg.fasta"), ctgs)
    base.write_fasta(os.path.join(d, "ref.fasta"), refs)
    subprocess.run([os.path.join(REFDIR, "kmer_counter"), "-t", "1", "-i", "reads.fq", "-o", "solid.bin", "-k", str(k)],
                   cwd=d, check=True, stdout=subprocess.DEVNULL)
    base.mecat(d, os.path.join(d, "reads.fq"), os.path.join(d, "ctg.fasta"), os.path.join(d, "r2c.ref"), nout=3)
    base.mecat(d, os.path.join(d, "reads.fq"), os.path.join(d, "ref.fasta"), os.path.join(d, "r2r.ref"), nout=3)
    recs = base.mecat(d, os.path.join(d, "ctg.fasta"), os.path.join(d, "ref.fasta"), os.path.join(d, "c2r.tmp"))
    os.remove(os.path.join(d, "c2r.tmp"))
    with open(os.path.join(d, "c2r.ref"), "w") as f:
        for h, a, b in recs:
            t = h.split("\t")
            t[0] = ctgs[int(t[0])][0]
            f.write("\t".join(t) + "\n" + a + "\n" + b + "\n")
    with open(os.path.join(d, "config.txt"), "w") as f:
        f.write("chrA\nreads.fq\nr2c.ref\nr2r.ref\nctg0\n1\nctg1\n1\nctg2\n0\n\n")


def run_travel(d, threads, out, eps=10, cov=2, min_len=50, graph="graph_ref.txt"):
    """The reference classes: graph build with ONE thread, PAssembly::testTravel5 with `threads` (oracle/ref_pagraph_dump.cpp)."""
    os.makedirs(os.path.join(d, out), exist_ok=True)
    r = subprocess.run([os.path.join(REFDIR, "pagraph_dump"), "1", "solid.bin", "ctg.fasta", "ref.fasta", ".", "c2r.ref", str(eps), str(cov),
                        graph, str(threads), out, str(min_len)], cwd=d, check=True, capture_output=True, text=True)
    return r.stdout


def run_binary(d, out, eps=10, cov=2):
    os.makedirs(os.path.join(d, out), exist_ok=True)
    subprocess.run([os.path.join(REFDIR, "pagraph"), "-t", "1", "-r", "dummy", "-k", "solid.bin", "-c", "ctg.fasta", "-R", "ref.fasta", "-p", ".",
                    "-a", "c2r.ref", "-o", out, "-r", "50", "--epsilon", str(eps), "-v", str(cov)], cwd=d, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def unpack(dest, name="pagraph_travel.tar.xz"):
    base.unpack(dest, name)


def main():
    kw = {}
    for a in sys.argv[1:]:
        k, v = a.split("=")
        kw[k] = int(v)
    d = tempfile.mkdtemp()
    build(d, **kw)
    log = run_travel(d, 1, "t1", graph="graph.txt")
    run_travel(d, 8, "t8")
    run_binary(d, "bin1")
    for n in sorted(os.listdir(os.path.join(d, "bin1"))):
        assert open(os.path.join(d, "bin1", n), "rb").read() == open(os.path.join(d, "t1", n), "rb").read(), n
    print("leaps:", log.count("leap"), "rounds:", log.count("choose"), "files:", sorted(os.listdir(os.path.join(d, "t1"))))
    print("t1 == t8:", all(open(os.path.join(d, "t1", n), "rb").read() == open(os.path.join(d, "t8", n), "rb").read()
                           for n in os.listdir(os.path.join(d, "t1"))))
    if "--dry" in os.environ.get("AG2_GEN", ""):
        print(d)
        return
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w") as tar:
        for n in base.FILES + ["graph.txt"]:
            tar.add(os.path.join(d, n), arcname=n)
        for t in ("t1", "t8"):
            for n in sorted(os.listdir(os.path.join(d, t))):
                tar.add(os.path.join(d, t, n), arcname=t + "/" + n)
    with open(os.path.join(HERE, "pagraph_travel.tar.xz"), "wb") as f:
        f.write(lzma.compress(buf.getvalue(), preset=9))
    print("->", os.path.getsize(os.path.join(HERE, "pagraph_travel.tar.xz")))
    print(d)


if __name__ == "__main__":
    main()
