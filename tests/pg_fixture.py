"""Helpers shared by the PAGraph traversal tests: the fixture's input set and its golden graph as host arrays."""
import os

import numpy as np


def read_fasta(path):
    names, seqs = [], []
    for line in open(path, "rb").read().split(b"\n"):
        if line.startswith(b">"):
            names.append(line[1:].split()[0].decode())
            seqs.append(b"")
        elif names:
            seqs[-1] += line
    return names, seqs


def read_blocks(path, ctg_names):
    """<pre dir>/config.txt -> per block [(contig index, forward flag)] (PGM/pagraph.cpp:29-49)."""
    cfg = open(path).read().split("\n")
    blocks, i = [], 0
    while i < len(cfg) and cfg[i]:
        i += 4
        use = []
        while i < len(cfg) and cfg[i]:
            use.append((ctg_names.index(cfg[i]), cfg[i + 1].strip() == "1"))
            i += 2
        i += 1
        blocks.append(use)
    return blocks


def graphs_from_dump(path, n_vertices):
    """graph.txt (the dump of the reference classes) -> one aligngraph2_b200.pagraph.Graph (CSR) per config block."""
    from aligngraph2_b200 import pagraph
    from oracle import binding
    out = []
    for cfg in binding.parse_graph_dump(open(path, "rb").read()):
        po, eo = np.zeros(n_vertices + 1, np.int64), np.zeros(n_vertices + 1, np.int64)
        for v, (_, pos, edges) in cfg.items():
            po[v + 1], eo[v + 1] = len(pos), len(edges)
        po, eo = np.cumsum(po), np.cumsum(eo)
        ctg, ref, cnt = np.zeros(po[-1], np.uint32), np.zeros(po[-1], np.uint32), np.zeros(po[-1], np.uint16)
        to, step = np.zeros(eo[-1], np.uint32), np.zeros(eo[-1], np.int32)
        for v, (_, pos, edges) in cfg.items():
            if pos:
                a = np.array(pos, np.int64)
                ctg[po[v]:po[v + 1]], ref[po[v]:po[v + 1]], cnt[po[v]:po[v + 1]] = a[:, 0], a[:, 1], a[:, 2]
            if edges:
                a = np.array(edges, np.int64)
                to[eo[v]:eo[v + 1]], step[eo[v]:eo[v + 1]] = a[:, 0], a[:, 1]
        out.append(pagraph.Graph(po, ctg, ref, cnt, eo, to, step))
    return out


def load_inputs(d):
    words = np.fromfile(os.path.join(d, "solid.bin"), np.uint64)
    codes = np.unique(words)           # the k-mer file read from byte 0: the leading k is one of the "k-mers" (FileKmerIterator)
    ctgs, refs = read_fasta(os.path.join(d, "ctg.fasta")), read_fasta(os.path.join(d, "ref.fasta"))
    return codes, int(words[0]), ctgs, refs, read_blocks(os.path.join(d, "config.txt"), ctgs[0])


def compare_dirs(got, want, skip=()):
    """-> list of differences between two output directories (byte for byte)."""
    bad = []
    for n in sorted(set(os.listdir(got)) | set(os.listdir(want))):
        if n in skip:
            continue
        a, b = os.path.join(got, n), os.path.join(want, n)
        if not os.path.exists(a) or not os.path.exists(b):
            bad.append(f"{n}: only in {'want' if os.path.exists(b) else 'got'}")
        elif open(a, "rb").read() != open(b, "rb").read():
            bad.append(f"{n}: differs")
    return bad
