"""Generates tests/golden/diff_blocks.npz and diff_go.npz from the UNMODIFIED vanilla-MECAT2 DiffAligner
(oracle/_ref/libref_mecat_vanilla.so = thirdparty/mecat/src/common/{defs,gapalign,diff_gapalign}.cpp behind
oracle/ref_diff_shim.cpp).

Run in the build container (needs /root/reference):  python tests/golden/gen_diff_golden.py
SURVEY row N2 ("next"): these fixtures pin oracle/ag2_diff.c, the checker a CUDA path for that row will be held against.
Blocks: the inputs of one `Align` call as dw_in_one_direction makes it (both directions; CLR-like pairs, unrelated pairs,
tandem repeats, blocks of a few bases) with end points, distance and both alignment strings.  Extensions: DiffAligner::go
on whole read / window pairs around a seed, incl. a seed at position 0 and at the end.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
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


def block_inputs(rng, kind):
    n = int(rng.integers(1, 720)) if kind != "tiny" else int(rng.integers(1, 12))
    t = rng.integers(0, 4, n).astype(np.uint8)
    if kind == "repeat":
        t = np.tile(rng.integers(0, 4, int(rng.integers(2, 9))).astype(np.uint8), n)[:n]
    q = rng.integers(0, 4, int(rng.integers(1, 600))).astype(np.uint8) if kind == "unrelated" else mutate(t, 0.15, rng)
    if len(q) == 0:
        q = t[:1].copy()
    return q, t


def main():
    binding.build(ref=True)
    ref = binding.DiffRef()
    rng = np.random.default_rng(20261017)
    out = {}
    kinds = ["clr"] * 28 + ["unrelated"] * 6 + ["repeat"] * 8 + ["tiny"] * 6
    for i, kind in enumerate(kinds):
        q, t = block_inputs(rng, kind)
        fwd = i & 1
        a = ref.block(q, t, fwd)
        out[f"q_{i}"], out[f"t_{i}"], out[f"fwd_{i}"] = q, t, np.int32(fwd)
        out[f"res_{i}"] = np.array([a["rc"], a["q_s"], a["q_e"], a["t_s"], a["t_e"], a["dist"], a["n"]], dtype=np.int32)
        out[f"qstr_{i}"] = np.frombuffer(a["qstr"], dtype=np.uint8)
        out[f"tstr_{i}"] = np.frombuffer(a["tstr"], dtype=np.uint8)
    out["n"] = np.int32(len(kinds))
    np.savez_compressed(os.path.join(HERE, "diff_blocks.npz"), **out)

    out = {}
    n_go = 14
    for i in range(n_go):
        n = int(rng.integers(300, 5000))
        t = rng.integers(0, 4, n).astype(np.uint8)
        q = mutate(t, 0.15, rng)
        pad = rng.integers(0, 4, int(rng.integers(0, 400))).astype(np.uint8)
        T = np.concatenate([pad, t, pad])
        ts = 0 if i == 0 else n if i == 1 else int(rng.integers(0, n + 1))
        qs = 0 if i == 0 else len(q) if i == 1 else min(len(q), int(ts * len(q) / max(1, n)))
        a = ref.go(q, qs, T, ts + len(pad), 0)
        out[f"q_{i}"], out[f"t_{i}"] = q, T
        out[f"res_{i}"] = np.array([qs, ts + len(pad), a["ok"], a["qoff"], a["qend"], a["toff"], a["tend"], a["aln_size"]], dtype=np.int32)
        out[f"qaln_{i}"] = np.frombuffer(a["qaln"], dtype=np.uint8)
        out[f"taln_{i}"] = np.frombuffer(a["taln"], dtype=np.uint8)
    out["n"] = np.int32(n_go)
    np.savez_compressed(os.path.join(HERE, "diff_go.npz"), **out)
    ref.close()
    print("wrote diff_blocks.npz, diff_go.npz")


if __name__ == "__main__":
    main()
