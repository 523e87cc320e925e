"""Fixture for the PAGraph A-Bruijn build and its golden graph dump from the UNMODIFIED
reference sources (oracle/_ref/pagraph_dump = src/tools/*.cpp + oracle/ref_pagraph_dump.cpp).

  python tests/golden/gen_pagraph_golden.py            # needs oracle/_ref/{mecat2ref,kmer_counter,pagraph_dump}

Writes tests/golden/pagraph_small.tar.xz: the pagraph input set
  reads.fq  ctg.fasta  ref.fasta  solid.bin  r2c.ref  r2r.ref  c2r.ref  config.txt
and graph.txt, the dump of every vertex (positions with counts, edges) after PositionProcessor::process().
The alignment files are what the reference's own aligner (oracle/_ref/mecat2ref, `-p` output passed through
script/filter.py's rule) writes for the synthetic reads / contigs, i.e. the files the pipeline hands to pagraph.

The data is built to reach: two references (PositionMapper offsets, the reference white list), a contig used in
reverse orientation (flag 0), contig bases without a contig->reference alignment ((0,0) entries), reads that align to
a contig but not the reference and the other way round, several alignments per read, reads shorter than k.
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
from aligngraph2_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
COMP = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")
FILES = ["reads.fq", "ctg.fasta", "ref.fasta", "solid.bin", "r2c.ref", "r2r.ref", "c2r.ref", "config.txt"]


def noisy(rng, seq, rate, ins=0.6, dele=0.25):
    out = bytearray()
    for c in seq:
        x = rng.random()
        if x < rate * ins:
            out.append(b"ACGT"[rng.integers(0, 4)])
            out.append(c)
        elif x < rate * (ins + dele):
            pass
        elif x < rate:
            out.append(b"ACGT"[(b"ACGT".index(c) + rng.integers(1, 4)) & 3])
        else:
            out.append(c)
    return bytes(out)


def rc(s):
    return s[::-1].translate(COMP)


def write_fasta(path, recs):
    with open(path, "wb") as f:
        for name, seq in recs:
            f.write(b">" + name.encode() + b"\n")
            for i in range(0, len(seq), 70):
                f.write(seq[i:i + 70] + b"\n")


def mecat(d, reads, ref, out, nout=1):
    """The reference aligner as AlignGraph2.py:265-277 runs it, then script/filter.py's rule (equal-length lines)."""
    wrk = tempfile.mkdtemp(dir=d)
    cmd = [os.path.join(REFDIR, "mecat2ref"), "-t", "1", "-d", os.path.abspath(reads), "-r", os.path.abspath(ref), "-b", str(nout),
           "-w", "./wrk", "-o", "o.txt", "-p", "p.txt", "-l", "0.5", "-u", "2.0", "-z", "200", "-y", "0.9"]
    subprocess.run(cmd, cwd=wrk, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    lines = open(os.path.join(wrk, "p.txt")).read().split("\n")
    recs = []
    for i in range(0, len(lines) - 2, 3):
        if len(lines[i + 1]) == len(lines[i + 2]):
            recs.append(lines[i:i + 3])
    shutil.rmtree(wrk)
    with open(out, "w") as f:
        for r in recs:
            f.write("\n".join(r) + "\n")
    return recs


def build(d, seed=11, genome_len=60000, n_reads=120, tlen=4000, k=9):
    rng = np.random.default_rng(seed)
    G = synth.make_reference(rng, genome_len).tobytes()
    cut = genome_len * 3 // 5
    # a 1.5 kb repeat in both halves: reads get alignments to two contigs / two references
    rep = G[5000:6500]
    G = G[:cut + 9000] + rep + G[cut + 10500:]
    refs = [("chrA", noisy(rng, G[:cut], 0.05, 0.3, 0.3)), ("chrB extra words", noisy(rng, G[cut:], 0.05, 0.3, 0.3))]
    ctgs = [("ctg0", G[1500:cut * 2 // 5]),
            ("ctg1", rc(G[cut * 2 // 5 + 2500:cut - 1000])),          # used in reverse orientation
            ("ctg2", G[cut + 2000:genome_len - 1500] + synth.make_reference(rng, 1200).tobytes())]  # tail aligns nowhere
    reads = []
    for i in range(n_reads):
        s = int(rng.integers(0, genome_len - tlen))
        rd = noisy(rng, G[s:s + tlen], 0.15)
        if i % 2:
            rd = rc(rd)
        if i % 31 == 5:
            b = bytearray(rd)
            for j in rng.integers(0, len(b), size=6):
                b[j] = ord("N")
            rd = bytes(b)
        reads.append(rd)
    reads.append(b"ACGTACG")                                        # shorter than k
    reads.append(synth.make_reference(rng, 2500).tobytes())          # aligns nowhere
    with open(os.path.join(d, "reads.fq"), "wb") as f:                # names = mecat2ref's 1-based running ids
        for i, rd in enumerate(reads):
            f.write(b"@%d\n" % (i + 1) + rd + b"\n+\n" + b"I" * len(rd) + b"\n")
    write_fasta(os.path.join(d, "ctg.fasta"), ctgs)
    write_fasta(os.path.join(d, "ref.fasta"), refs)
    subprocess.run([os.path.join(REFDIR, "kmer_counter"), "-t", "1", "-i", "reads.fq", "-o", "solid.bin", "-k", str(k)],
                   cwd=d, check=True, stdout=subprocess.DEVNULL)
    mecat(d, os.path.join(d, "reads.fq"), os.path.join(d, "ctg.fasta"), os.path.join(d, "r2c.ref"), nout=3)
    mecat(d, os.path.join(d, "reads.fq"), os.path.join(d, "ref.fasta"), os.path.join(d, "r2r.ref"), nout=3)
    # contig -> reference: the pipeline gets these from MUMmer via paf2aln; same 3-line format.  The aligner numbers
    # FASTA queries 0,1,2..: put the contig names back.
    recs = mecat(d, os.path.join(d, "ctg.fasta"), os.path.join(d, "ref.fasta"), os.path.join(d, "c2r.tmp"))
    os.remove(os.path.join(d, "c2r.tmp"))
    with open(os.path.join(d, "c2r.ref"), "w") as f:
        for h, a, b in recs:
            t = h.split("\t")
            t[0] = ctgs[int(t[0])][0]
            f.write("\t".join(t) + "\n" + a + "\n" + b + "\n")
    with open(os.path.join(d, "config.txt"), "w") as f:
        f.write("chrA\nreads.fq\nr2c.ref\nr2r.ref\nctg0\n1\nctg1\n0\n\n")
        f.write("chrB\nreads.fq\nr2c.ref\nr2r.ref\nctg2\n1\n\n")


def run_dump(d, eps=10, cov=2, threads=1, out="graph.txt"):
    subprocess.run([os.path.join(REFDIR, "pagraph_dump"), str(threads), "solid.bin", "ctg.fasta", "ref.fasta", ".", "c2r.ref",
                    str(eps), str(cov), out], cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return open(os.path.join(d, out), "rb").read()


def unpack(dest, name="pagraph_small.tar.xz"):
    with lzma.open(os.path.join(HERE, name)) as xz, tarfile.open(fileobj=xz) as tar:
        tar.extractall(dest, filter="data")


def main():
    d = tempfile.mkdtemp()
    build(d)
    run_dump(d)
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w") as tar:
        for n in FILES + ["graph.txt"]:
            tar.add(os.path.join(d, n), arcname=n)
    with open(os.path.join(HERE, "pagraph_small.tar.xz"), "wb") as f:
        f.write(lzma.compress(buf.getvalue(), preset=9))
    for n in FILES + ["graph.txt"]:
        print(n, os.path.getsize(os.path.join(d, n)))
    print("->", os.path.getsize(os.path.join(HERE, "pagraph_small.tar.xz")))
    print(d)


if __name__ == "__main__":
    main()
