"""Deterministic synthetic PacBio-CLR-like inputs.

Reference = i.i.d. uniform ACGT.  Read i = a 10 kb template cut from the
reference, odd i reverse-complemented, then per template base one draw r:
r < 0.09 -> insert a uniform base before it, r < 0.1275 -> delete it,
r < 0.15 -> substitute it, else copy (15 % error: 60 % ins / 25 % del /
15 % sub).  numpy is the only dependency so fixtures can be regenerated on
any box from (seed, sizes); the generator is not part of the parity contract
(the reference binary and this package are always run on the same files).
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[:] = np.arange(256, dtype=np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP[_a] = _b

P_INS, P_DEL, P_SUB = 0.09, 0.0375, 0.0225


def make_reference(rng: np.random.Generator, length: int) -> np.ndarray:
    """Uniform ACGT as uint8 ASCII."""
    return _ACGT[rng.integers(0, 4, size=length, dtype=np.uint8)]


def revcomp(seq: np.ndarray) -> np.ndarray:
    return _COMP[seq[::-1]]


def make_read(rng: np.random.Generator, ref: np.ndarray, tlen: int, reverse: bool):
    """One read; returns (ascii uint8 array, template start, ops array).

    The returned sequence is in *read* orientation (already reverse-complemented
    when ``reverse``); ``ops`` is per template base: 0 copy, 1 ins-before,
    2 del, 3 sub.
    """
    start = int(rng.integers(0, len(ref) - tlen))
    tpl = ref[start:start + tlen]
    r = rng.random(tlen)
    ops = np.zeros(tlen, dtype=np.uint8)
    ops[r < P_INS + P_DEL + P_SUB] = 3
    ops[r < P_INS + P_DEL] = 2
    ops[r < P_INS] = 1
    # output count per template base: ins -> 2 (random base + the base), del -> 0
    cnt = np.ones(tlen, dtype=np.int64)
    cnt[ops == 1] = 2
    cnt[ops == 2] = 0
    off = np.concatenate(([0], np.cumsum(cnt)))
    out = np.empty(int(off[-1]), dtype=np.uint8)
    keep = ops != 2
    # position of the template-derived base (last slot of its group)
    pos_main = off[1:][keep] - 1
    main = tpl[keep].copy()
    sub_mask = (ops == 3)[keep]
    nsub = int(sub_mask.sum())
    if nsub:
        # substitute with one of the three other bases, uniformly
        code = np.searchsorted(_ACGT, main[sub_mask])
        code = (code + rng.integers(1, 4, size=nsub)) & 3
        main[sub_mask] = _ACGT[code]
    out[pos_main] = main
    ins_pos = off[:-1][ops == 1]
    out[ins_pos] = _ACGT[rng.integers(0, 4, size=len(ins_pos), dtype=np.uint8)]
    if reverse:
        out = revcomp(out)
    return out, start, ops


def make_dataset(seed: int, ref_len: int, n_reads: int, tlen: int = 10000):
    """Returns (ref ascii uint8, list of read ascii uint8, truth list)."""
    rng = np.random.default_rng(seed)
    ref = make_reference(rng, ref_len)
    reads, truth = [], []
    for i in range(n_reads):
        rd, start, _ = make_read(rng, ref, tlen, reverse=bool(i & 1))
        reads.append(rd)
        truth.append((start, i & 1))
    return ref, reads, truth


def write_fasta(path: str, name: str, seq: np.ndarray, width: int = 80) -> None:
    with open(path, "wb") as f:
        f.write(b">" + name.encode() + b"\n")
        for i in range(0, len(seq), width):
            f.write(seq[i:i + width].tobytes())
            f.write(b"\n")


def write_fastq(path: str, reads) -> None:
    with open(path, "wb") as f:
        for i, rd in enumerate(reads):
            f.write(b"@read%d\n" % i)
            f.write(rd.tobytes())
            f.write(b"\n+\n")
            f.write(b"I" * len(rd))
            f.write(b"\n")


def spike_repeats(rng: np.random.Generator, ref: np.ndarray, unit_len: int, copies: int) -> np.ndarray:
    """Overwrite ``copies`` random windows of ``ref`` with one repeat unit
    (exercises the >128 bucket mask and the alpha/beta vote)."""
    ref = ref.copy()
    unit = make_reference(rng, unit_len)
    for _ in range(copies):
        s = int(rng.integers(0, len(ref) - unit_len))
        ref[s:s + unit_len] = unit
    return ref


if __name__ == "__main__":
    import argparse
    import os

    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--seed", type=int, default=20261017)
    ap.add_argument("--ref-len", type=int, default=1_000_000)
    ap.add_argument("--reads", type=int, default=1000)
    ap.add_argument("--tlen", type=int, default=10000)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    ref, reads, _ = make_dataset(a.seed, a.ref_len, a.reads, a.tlen)
    write_fasta(os.path.join(a.out, "ref.fa"), "chr1", ref)
    write_fastq(os.path.join(a.out, "reads.fq"), reads)


# ---------------------------------------------------------------------------------------------
# Batched generator (torch; runs on the GPU for bench-sized inputs, on the CPU for tests).
# Same per-base rule as make_read; additionally returns one exact-match anchor per read, i.e. the
# (loc1, loc2) a seed hit would hand to extend_candidate (mecat2ref_aux.cpp:226-227).
# ---------------------------------------------------------------------------------------------
def make_batch_torch(seed: int, ref_len: int, n_reads: int, tlen: int = 10000, device="cpu",
                     chunk_reads: int = 8192, ref=None):
    """Returns dict of torch tensors on ``device``:
    ref (uint8 ASCII [ref_len]), bases (uint8 ASCII, reads concatenated in FILE orientation: odd
    reads reverse-complemented), offsets (int64 [n+1]), strand (int32 [n]: 1 for odd reads),
    loc1 (int64, 1-based reference position of the anchor), loc2 (int32, position in the ORIENTED
    read), start (int64 template start)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    if ref is None:
        ref_codes = torch.randint(0, 4, (ref_len,), generator=g, device=device, dtype=torch.uint8)
    else:
        lut = torch.zeros(256, dtype=torch.uint8, device=device)
        for i, ch in enumerate(b"ACGT"):
            lut[ch] = i
        ref_codes = lut[torch.as_tensor(ref, device=device).long()]
        ref_len = ref_codes.numel()
    out_chunks, len_chunks, loc1_c, loc2_c, start_c = [], [], [], [], []
    ar = torch.arange(tlen, device=device)
    K = 13
    for lo in range(0, n_reads, chunk_reads):
        m = min(chunk_reads, n_reads - lo)
        start = torch.randint(0, ref_len - tlen, (m,), generator=g, device=device)
        r = torch.rand((m, tlen), generator=g, device=device)
        ins = r < P_INS
        dele = (~ins) & (r < P_INS + P_DEL)
        sub = (~ins) & (~dele) & (r < P_INS + P_DEL + P_SUB)
        cnt = 1 + ins.to(torch.int32) - dele.to(torch.int32)
        off = torch.cumsum(cnt, dim=1)                      # inclusive: off[:, p] = end of base p's group
        lens = off[:, -1].to(torch.int64)
        tpl = ref_codes[(start[:, None] + ar[None, :])]
        main = torch.where(sub, (tpl + torch.randint(1, 4, (m, tlen), generator=g, device=device, dtype=torch.uint8)) & 3, tpl)
        insb = torch.randint(0, 4, (m, tlen), generator=g, device=device, dtype=torch.uint8)
        roff = torch.zeros(m + 1, dtype=torch.int64, device=device)
        roff[1:] = torch.cumsum(lens, 0)
        flat = torch.empty(int(roff[-1]), dtype=torch.uint8, device=device)
        odd = ((torch.arange(m, device=device) + lo) & 1).bool()
        # forward-orientation position of each emitted base, then file orientation for odd reads
        pos_main = (off - 1).to(torch.int64)
        pos_ins = (off - 2).to(torch.int64)

        def place(pos, codes, mask):
            p = torch.where(odd[:, None], lens[:, None] - 1 - pos, pos) + roff[:-1, None]
            c = torch.where(odd[:, None], 3 - codes, codes)
            flat[p[mask]] = acgt[c[mask].long()]

        place(pos_main, main, ~dele)
        place(pos_ins, insb, ins)
        # anchor: K consecutive plain copies, nearest to a random template position
        bad = (ins | dele | sub).to(torch.int32)
        cs = torch.cumsum(bad, dim=1)
        win = cs[:, K - 1:] - torch.cat([torch.zeros((m, 1), dtype=cs.dtype, device=device), cs[:, :-K]], dim=1)
        ok = win == 0                                       # [m, tlen-K+1]: bases p..p+K-1 are copies
        u = torch.randint(0, tlen - K, (m, 1), generator=g, device=device)
        dist = torch.where(ok, (ar[None, :tlen - K + 1] - u).abs(), torch.full_like(u, 1 << 30))
        p = torch.argmin(dist, dim=1)
        loc2 = torch.gather(pos_main, 1, p[:, None])[:, 0]
        loc1_c.append(start + p + 1)
        loc2_c.append(loc2.to(torch.int32))
        start_c.append(start)
        out_chunks.append(flat)
        len_chunks.append(lens)
    lens = torch.cat(len_chunks)
    offsets = torch.zeros(n_reads + 1, dtype=torch.int64, device=device)
    offsets[1:] = torch.cumsum(lens, 0)
    return dict(ref=acgt[ref_codes.long()], bases=torch.cat(out_chunks), offsets=offsets,
                strand=(torch.arange(n_reads, device=device) & 1).to(torch.int32),
                loc1=torch.cat(loc1_c), loc2=torch.cat(loc2_c), start=torch.cat(start_c))


_COMP_TABLE = bytes.maketrans(b"ACGT", b"TGCA")


def orient(read: bytes, strand: int) -> bytes:
    """The read as reference_mapping() hands it to extend_candidate: as given for 'F', reversed and
    with upper-case ACGT complemented for 'R' (impl_large.cpp:799-833)."""
    return read if not strand else read[::-1].translate(_COMP_TABLE)
