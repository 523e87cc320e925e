"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libref_mecat.so).

Run in the build container (needs /root/reference):  python tests/golden/gen_golden.py
The reference ships no tests or golden vectors for this path, so these
fixtures -- outputs of the reference's own xdrop_align / XdropAligner::go / extend_candidate on
seeded synthetic inputs -- are what pins oracle/ag2_oracle.c and, through it, the CUDA path.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from aligngraph2_b200 import synth  # noqa: E402
from oracle import binding  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def mutate(seq, rate, rng):
    out = []
    for c in seq:
        x = rng.random()
        if x < rate * 0.6:
            out.append(int(rng.integers(0, 4)))
            out.append(int(c))
        elif x < rate * 0.85:
            pass
        elif x < rate:
            out.append(int((c + rng.integers(1, 4)) & 3))
        else:
            out.append(int(c))
    return np.array(out, dtype=np.uint8)


def gen_blocks(ref, n=48, seed=11):
    rng = np.random.default_rng(seed)
    rows = []
    while len(rows) < n:
        M = int(rng.integers(1, 719))
        rate = float(rng.choice([0.0, 0.05, 0.15, 0.3, 0.5]))
        A = rng.integers(0, 4, size=M).astype(np.uint8)
        if rng.random() < 0.15:  # short tandem repeats: wide bands, many ties
            A = np.tile(rng.integers(0, 4, size=int(rng.integers(1, 6))), M)[:M].astype(np.uint8)
        B = mutate(A, rate, rng)
        if rng.random() < 0.1:
            B = rng.integers(0, 4, size=int(rng.integers(1, 719))).astype(np.uint8)
        N = min(len(B), 718)
        if N == 0:
            continue
        B = B[:N]
        fwd = bool(rng.integers(0, 2))
        if not fwd:
            A, B = A[::-1].copy(), B[::-1].copy()
        s, ae, be, ops = ref.block(A, M, B, N, fwd)
        rows.append(dict(A=A, B=B, fwd=fwd, score=s, ae=ae, be=be, ops=ops))
    np.savez_compressed(os.path.join(HERE, "xdrop_blocks.npz"),
                        n=len(rows),
                        **{f"{k}_{i}": np.asarray(r[k]) for i, r in enumerate(rows) for k in r})


def gen_extend(ref, seed=20261017):
    d = synth.make_batch_torch(seed, 120_000, 16, 3000)
    refb = d["ref"].numpy().tobytes()
    bases = d["bases"].numpy()
    off = d["offsets"].numpy()
    rng = np.random.default_rng(seed)
    bases = bases.copy()
    # soft-masked and N bases in two reads: the reverse strand must not complement them
    for r in (2, 5):
        seg = bases[off[r]:off[r + 1]]
        idx = rng.integers(0, len(seg), size=30)
        seg[idx[:20]] |= 0x20
        seg[idx[20:]] = ord("N")
    recs = []
    for i in range(16):
        rd = bases[off[i]:off[i + 1]].tobytes()
        s = int(d["strand"][i])
        loc1, loc2 = int(d["loc1"][i]), int(d["loc2"][i])
        if i == 7:   # a seed at the very start of the read: empty left extension
            loc1, loc2 = loc1 - loc2 // 1, 0
            # keep loc1 consistent only approximately; the aligner does not need an exact seed
            loc1 = max(1, int(d["start"][i]) + 1)
        r = ref.extend(refb, synth.orient(rd, s), loc1, loc2)
        recs.append(dict(strand=s, loc1=loc1, loc2=loc2, ok=r["ok"], qb=r.get("qb", 0), qe=r.get("qe", 0),
                         sb=r.get("sb", 0), se=r.get("se", 0),
                         qaln=np.frombuffer(r.get("qaln", b""), dtype=np.uint8),
                         taln=np.frombuffer(r.get("taln", b""), dtype=np.uint8)))
    np.savez_compressed(os.path.join(HERE, "extend_candidates.npz"), ref=np.frombuffer(refb, dtype=np.uint8),
                        bases=bases, offsets=off, n=len(recs),
                        **{f"{k}_{i}": np.asarray(r[k]) for i, r in enumerate(recs) for k in r})


if __name__ == "__main__":
    binding.build()
    ref = binding.RefLib()
    gen_blocks(ref)
    gen_extend(ref)
    print("golden vectors written to", HERE)
