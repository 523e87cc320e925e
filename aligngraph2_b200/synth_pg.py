"""Synthetic `pagraph` input sets of any size (BASELINE configs[3]: the A-Bruijn build on pre-aligned reads).

The reference pipeline gets its three `.ref` alignment files from aligners (mecat2ref+, vanilla mecat2ref, MUMmer); at
bench sizes that is hours of CPU, so this generator writes the alignments it KNOWS instead: every read is a noisy copy of
a genome window (the CLR error profile of synth.py), the contigs are genome windows, the reference is a 5 %-diverged copy
of the genome with recorded edits, and the alignment columns are composed from those edit records.  The files are
ordinary pagraph inputs (the CPU checker and the reference binary read them too); the generator is not part of the parity
contract.
"""
from __future__ import annotations

import os

import numpy as np

from . import synth

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_DASH = ord("-")


def _edit(rng, n, p_ins, p_del, p_sub):
    """per base: 0 copy, 1 insert-before, 2 delete, 3 substitute"""
    r = rng.random(n)
    ops = np.zeros(n, dtype=np.uint8)
    ops[r < p_ins + p_del + p_sub] = 3
    ops[r < p_ins + p_del] = 2
    ops[r < p_ins] = 1
    return ops


def _apply(rng, codes, ops):
    """(main base per position, inserted base per position) as 0..3 codes"""
    main = codes.copy()
    sub = ops == 3
    main[sub] = (main[sub] + rng.integers(1, 4, size=int(sub.sum()), dtype=np.uint8)) & 3
    ins = rng.integers(0, 4, size=len(codes), dtype=np.uint8)
    return main, ins


def _columns(q_ins, q_main, q_has, t_ins, t_main, t_has, q_ins_mask, t_ins_mask):
    """Compose alignment columns per template base: [query insert] [target insert] [main column]."""
    n = len(q_main)
    q = np.full((n, 3), _DASH, dtype=np.uint8)
    t = np.full((n, 3), _DASH, dtype=np.uint8)
    ok = np.zeros((n, 3), dtype=bool)
    q[:, 0] = _ACGT[q_ins]
    ok[:, 0] = q_ins_mask
    t[:, 1] = _ACGT[t_ins]
    ok[:, 1] = t_ins_mask
    q[q_has, 2] = _ACGT[q_main[q_has]]
    t[t_has, 2] = _ACGT[t_main[t_has]]
    ok[:, 2] = q_has | t_has
    return q[ok], t[ok]


def make_input_set(out_dir: str, seed: int, genome_len: int, n_reads: int, tlen: int = 10000, n_ctg: int = 8,
                   divergence: float = 0.05, solid_words: np.ndarray | None = None) -> dict:
    """Writes reads.fq ctg.fasta ref.fasta r2c.ref r2r.ref c2r.ref config.txt (and solid.bin if solid_words is given)
    into out_dir; returns sizes.  Reads are named 1..n (mecat2ref's running ids)."""
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(seed)
    G = rng.integers(0, 4, size=genome_len, dtype=np.uint8)
    # reference = diverged genome (ins 0.3 / del 0.3 / sub 0.4 of the divergence)
    r_ops = _edit(rng, genome_len, divergence * 0.3, divergence * 0.3, divergence * 0.4)
    r_main, r_insb = _apply(rng, G, r_ops)
    r_cnt = np.ones(genome_len, dtype=np.int64)
    r_cnt[r_ops == 1] = 2
    r_cnt[r_ops == 2] = 0
    r_end = np.cumsum(r_cnt)                      # reference coordinate after the group of genome base g
    r_start = r_end - r_cnt
    R = np.empty(int(r_end[-1]), dtype=np.uint8)
    keep = r_ops != 2
    R[r_end[keep] - 1] = r_main[keep]
    R[r_start[r_ops == 1]] = r_insb[r_ops == 1]
    # contigs = genome windows separated by gaps
    span = genome_len // n_ctg
    gap = min(6000, span // 10)
    ctgs = [(f"ctg{i}", i * span + gap // 2, (i + 1) * span - gap // 2) for i in range(n_ctg)]
    with open(os.path.join(out_dir, "ctg.fasta"), "wb") as f:
        for name, a, b in ctgs:
            f.write(b">" + name.encode() + b"\n")
            s = _ACGT[G[a:b]].tobytes()
            for i in range(0, len(s), 70):
                f.write(s[i:i + 70] + b"\n")
    synth.write_fasta(os.path.join(out_dir, "ref.fasta"), "chr1", _ACGT[R], 70)
    # contig -> reference (the file MUMmer + paf2aln produce in the pipeline)
    with open(os.path.join(out_dir, "c2r.ref"), "wb") as f:
        for name, a, b in ctgs:
            sl = slice(a, b)
            has_t = r_ops[sl] != 2
            q, t = _columns(np.zeros(b - a, np.uint8), G[sl], np.ones(b - a, bool), r_insb[sl], r_main[sl], has_t,
                            np.zeros(b - a, bool), r_ops[sl] == 1)
            f.write(b"%s\tchr1\tF\t%d\t0\t%d\t%d\t%d\t%d\t%d\n" % (name.encode(), b - a, b - a, b - a, r_start[a], r_end[b - 1], len(R)))
            f.write(q.tobytes() + b"\n" + t.tobytes() + b"\n")
    # reads and their two alignment files
    total_bases = columns = 0
    fq = open(os.path.join(out_dir, "reads.fq"), "wb")
    r2c = open(os.path.join(out_dir, "r2c.ref"), "wb")
    r2r = open(os.path.join(out_dir, "r2r.ref"), "wb")
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    for i in range(n_reads):
        ci = int(rng.integers(0, n_ctg))
        name, a, b = ctgs[ci]
        s = int(rng.integers(a, b - tlen - 1))
        sl = slice(s, s + tlen)
        ops = _edit(rng, tlen, synth.P_INS, synth.P_DEL, synth.P_SUB)
        main, insb = _apply(rng, G[sl], ops)
        has_q = ops != 2
        q_c, t_c = _columns(insb, main, has_q, np.zeros(tlen, np.uint8), G[sl], np.ones(tlen, bool), ops == 1, np.zeros(tlen, bool))
        q_r, t_r = _columns(insb, main, has_q, r_insb[sl], r_main[sl], r_ops[sl] != 2, ops == 1, r_ops[sl] == 1)
        read_fwd = q_c[q_c != _DASH].tobytes()
        n = len(read_fwd)
        rev = bool(i & 1)
        rd = read_fwd[::-1].translate(comp) if rev else read_fwd
        fq.write(b"@%d\n" % (i + 1) + rd + b"\n+\n" + b"I" * n + b"\n")
        strand = b"R" if rev else b"F"
        score = int((ops == 0).sum() // 100)
        r2c.write(b"%d\t%s\t%s\t%d\t0\t%d\t%d\t%d\t%d\t%d\n" % (i + 1, name.encode(), strand, score, n, n, s - a, s - a + tlen, b - a))
        r2c.write(q_c.tobytes() + b"\n" + t_c.tobytes() + b"\n")
        r2r.write(b"%d\tchr1\t%s\t%d\t0\t%d\t%d\t%d\t%d\t%d\n" % (i + 1, strand, score, n, n, r_start[s], r_end[s + tlen - 1], len(R)))
        r2r.write(q_r.tobytes() + b"\n" + t_r.tobytes() + b"\n")
        total_bases += n
        columns += len(q_c) + len(q_r)
    for f in (fq, r2c, r2r):
        f.close()
    with open(os.path.join(out_dir, "config.txt"), "w") as f:
        f.write("chr1\nreads.fq\nr2c.ref\nr2r.ref\n" + "".join(f"{n}\n1\n" for n, _, _ in ctgs) + "\n")
    if solid_words is not None:
        np.asarray(solid_words, dtype=np.uint64).tofile(os.path.join(out_dir, "solid.bin"))
    return dict(read_bases=total_bases, columns=columns, ref_len=len(R), n_reads=n_reads, n_ctg=n_ctg)


def solid_words_from_reads(out_dir: str, k: int, threshold: float = 0.2, device: int = 0) -> np.ndarray:
    """solid_kmer_set.bin content for the reads of an input set, computed by this package's kmer_counter kernels
    (SURVEY row B1): uint64 k followed by the solid codes."""
    from .mecat2ref import Mecat2RefDevice

    reads, offs = [], [0]
    with open(os.path.join(out_dir, "reads.fq"), "rb") as f:
        for i, line in enumerate(f):
            if i % 4 == 1:
                reads.append(np.frombuffer(line.rstrip(b"\n"), dtype=np.uint8))
                offs.append(offs[-1] + len(reads[-1]))
    dev = Mecat2RefDevice(device)
    dev.load_reads(bases=np.concatenate(reads), offsets=np.asarray(offs, dtype=np.int64))
    codes, _cut = dev.solid_kmers(k, threshold)
    dev.close()
    words = np.concatenate((np.array([k], dtype=np.uint64), np.asarray(codes, dtype=np.uint64)))
    words.tofile(os.path.join(out_dir, "solid.bin"))
    return words
