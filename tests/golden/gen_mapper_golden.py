"""Stress fixture for the whole mecat2ref+ per-read path (index, votes, seeding, candidates, extension,
rescue, second pass) and its golden `.r` thread file from the UNMODIFIED reference binary.

  python tests/golden/gen_mapper_golden.py            # needs oracle/_ref/mecat2ref (build container)

Writes tests/golden/mapper_stress.npz (inputs) and tests/golden/mapper_stress_ref.tar.xz: what
`mecat2ref -t 1 ... -z 200` wrote -- <wrk>/1.r (thread file), <wrk>/chrindex.txt, the -o and -p outputs -- plus
the sha256 of <wrk>/0.fq and <wrk>/ref.fq (mapper_stress_ref.json).  The inputs are built to reach the branches a uniform
random genome never does: k-mer buckets above the 128 mask, similarity votes != 1, more than 20 seeds
per 1000-bp block (insert_loc), several candidates per read, chimeric reads (rescue_clipped_align),
unalignable reads (second pass), short reads, N and lower-case bases.
"""
import hashlib
import io
import json
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
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "mecat2ref")
ARGS = dict(b=1, l=0.5, u=2.0, z=200, y=0.9)


def noisy(rng, seq, rate):
    out = bytearray()
    for c in seq:
        x = rng.random()
        if x < rate * 0.6:
            out.append(b"ACGT"[rng.integers(0, 4)])
            out.append(c)
        elif x < rate * 0.85:
            pass
        elif x < rate:
            out.append(b"ACGT"[(b"ACGT".index(c) + rng.integers(1, 4)) & 3] if c in b"ACGT" else c)
        else:
            out.append(c)
    return bytes(out)


def build(seed=7):
    rng = np.random.default_rng(seed)
    parts = []
    unit = synth.make_reference(rng, 4000).tobytes()
    for i in range(60):
        parts.append(synth.make_reference(rng, int(rng.integers(2000, 9000))).tobytes())
        if i % 2 == 0:
            parts.append(noisy(rng, unit, 0.01))           # dispersed 4 kb repeat, 1 % diverged copies
        if i % 9 == 4:
            parts.append(b"AC" * 400 + b"A" * 300)          # low complexity: buckets far above the mask
        if i % 13 == 6:
            parts.append(b"N" * 150)
        if i == 30:
            parts.append(unit[:1500] * 6)                    # tandem copies
    chr1 = b"".join(parts)
    chr2 = synth.make_reference(rng, 60000).tobytes() + unit[::-1] + synth.make_reference(rng, 20000).tobytes().lower()
    chroms = [("chr1 first", chr1), ("chr2", chr2)]
    genome = b"".join(c for _, c in chroms).upper()
    reads = []
    comp = bytes.maketrans(b"ACGT", b"TGCA")

    def cut(length):
        s = int(rng.integers(0, len(genome) - length))
        return genome[s:s + length]

    for i in range(90):
        kind = i % 9
        if kind in (0, 1, 2):
            rd = noisy(rng, cut(int(rng.integers(1500, 9000))), 0.15)
        elif kind == 3:
            rd = noisy(rng, cut(int(rng.integers(3000, 8000))), 0.04)       # dense seeds: insert_loc
        elif kind == 4:
            rd = noisy(rng, cut(5000), 0.12) + noisy(rng, cut(4000), 0.12)  # chimera: clipped alignments
        elif kind == 5:
            a = cut(9000)
            rd = noisy(rng, a[:3500] + a[5500:], 0.12)                      # 2 kb deletion
        elif kind == 6:
            rd = synth.make_reference(rng, int(rng.integers(800, 4000))).tobytes()  # unalignable: second pass
        elif kind == 7:
            rd = noisy(rng, cut(int(rng.integers(300, 1400))), 0.1)         # short
        else:
            b = bytearray(noisy(rng, cut(6000), 0.13))
            for k in rng.integers(0, len(b), size=12):
                b[k] = ord("N")
            for k in rng.integers(0, len(b), size=40):
                b[k] |= 0x20
            rd = bytes(b)
        if rng.random() < 0.5:
            rd = rd[::-1].translate(comp)
        reads.append(rd)
    return chroms, reads


def write_inputs(d, chroms, reads):
    with open(os.path.join(d, "ref.fa"), "wb") as f:
        for name, seq in chroms:
            f.write(b">" + name.encode() + b"\n")
            for i in range(0, len(seq), 70):
                f.write(seq[i:i + 70] + b"\n")
    with open(os.path.join(d, "reads.fq"), "wb") as f:
        for i, rd in enumerate(reads):
            f.write(b"@r%d\n" % i + rd + b"\n+\n" + b"I" * len(rd) + b"\n")


def run_reference(d, threads=1):
    cmd = [REF_BIN, "-t", str(threads), "-d", "reads.fq", "-r", "ref.fa", "-b", str(ARGS["b"]), "-w", "./wrk", "-o", "o.txt",
           "-p", "p.txt", "-l", str(ARGS["l"]), "-u", str(ARGS["u"]), "-z", str(ARGS["z"]), "-y", str(ARGS["y"])]
    subprocess.run(cmd, cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return open(os.path.join(d, "wrk", "1.r"), "rb").read()


REF_FILES = ("wrk/1.r", "wrk/chrindex.txt", "o.txt", "p.txt")


def collect_reference_outputs(d):
    files = {n: open(os.path.join(d, n), "rb").read() for n in REF_FILES}
    hashes = {n: hashlib.sha256(open(os.path.join(d, n), "rb").read()).hexdigest() for n in ("wrk/0.fq", "wrk/ref.fq")}
    return files, hashes


if __name__ == "__main__":
    chroms, reads = build()
    d = tempfile.mkdtemp(prefix="m2r_golden_")
    try:
        write_inputs(d, chroms, reads)
        r = run_reference(d)
        files, hashes = collect_reference_outputs(d)
    finally:
        keep = os.environ.get("KEEP")
        if not keep:
            shutil.rmtree(d, ignore_errors=True)
        else:
            print("kept", d)
    offs = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum([len(x) for x in reads], out=offs[1:])
    np.savez_compressed(os.path.join(HERE, "mapper_stress.npz"),
                        chrom_names=np.array([n for n, _ in chroms]),
                        chrom_lens=np.array([len(s) for _, s in chroms]),
                        genome=np.frombuffer(b"".join(s for _, s in chroms), dtype=np.uint8),
                        bases=np.frombuffer(b"".join(reads), dtype=np.uint8), offsets=offs)
    with lzma.open(os.path.join(HERE, "mapper_stress_ref.tar.xz"), "wb", preset=9 | lzma.PRESET_EXTREME) as xz:
        with tarfile.open(fileobj=xz, mode="w") as tar:
            for name in REF_FILES:
                ti = tarfile.TarInfo(name)
                ti.size = len(files[name])
                tar.addfile(ti, io.BytesIO(files[name]))
    json.dump(hashes, open(os.path.join(HERE, "mapper_stress_ref.json"), "w"), indent=1)
    print("records:", r.count(b"\n") // 3, "bytes:", len(r))
