"""Golden for the whole per-read path at BASELINE configs[2]'s reference size: the UNMODIFIED reference binary
(`oracle/_ref/mecat2ref -t 1 ... -z 200`) on reads against a seeded 250 Mb uniform reference -- the size where a 13-mer
bucket holds 3.7 positions, a read strand has ~2 600 index hits and the block tables are no longer small.

  python tests/golden/gen_map250_golden.py            # needs oracle/_ref/mecat2ref (build container); ~10 min, 3 GB in /tmp

Nothing big is committed: reference and reads are regenerated from numpy seeds (`inputs()`), the golden
(tests/golden/map250_ref.json) holds the sha256 of <wrk>/1.r, of the -o / -p files, and per record the header line with
a short digest of its two alignment strings (so a mismatch names the read).
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from aligngraph2_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "mecat2ref")
SEED, REF_LEN, N_READS = 20261017, 250_000_000, 320
GOLDEN_JSON = os.path.join(HERE, "map250_ref.json")


def inputs(seed=SEED, ref_len=REF_LEN, n_reads=N_READS):
    """(reference ASCII uint8 [ref_len], list of read bytes).  Reads: CLR templates (10 kb, 15 % error) of both strands,
    every 16th a chimera of two places (rescue_clipped_align), every 16th + 8 unrelated to the reference (second pass),
    every 16th + 4 a template with 2.5 kb cut out of its middle (two clipped alignments that rescue links), every 16th + 12 a
    short one."""
    rng = np.random.default_rng(seed)
    ref = synth.make_reference(rng, ref_len)
    reads = []
    for i in range(n_reads):
        kind = i % 16
        if kind == 0:
            a, _, _ = synth.make_read(rng, ref, 5500, reverse=False)
            b, _, _ = synth.make_read(rng, ref, 4500, reverse=False)
            rd = np.concatenate([a, b])
            if i & 16:
                rd = synth.revcomp(rd)
        elif kind == 4:
            s0 = int(rng.integers(0, ref_len - 12000))
            tpl = np.concatenate([ref[s0:s0 + 4500], ref[s0 + 7000:s0 + 11500]])
            rd, _, _ = synth.make_read(rng, tpl, len(tpl) - 1, reverse=bool(i & 16))
        elif kind == 8:
            rd = synth.make_reference(rng, int(rng.integers(2000, 9000)))
        elif kind == 12:
            rd, _, _ = synth.make_read(rng, ref, int(rng.integers(1200, 3000)), reverse=bool(i & 16))
        else:
            rd, _, _ = synth.make_read(rng, ref, 10000, reverse=bool(i & 1))
        reads.append(rd.tobytes())
    return ref, reads


def record_digest(thread_file: bytes):
    """[header line + ' ' + sha1(qmap \\n smap)[:12]] per 3-line record"""
    lines = thread_file.split(b"\n")
    out = []
    for k in range(0, len(lines) - 2, 3):
        out.append(lines[k].decode() + " " + hashlib.sha1(lines[k + 1] + b"\n" + lines[k + 2]).hexdigest()[:12])
    return out


def run_reference(d, threads=1, args=("-b", "1", "-l", "0.5", "-u", "2.0", "-z", "200", "-y", "0.9")):
    cmd = [REF_BIN, "-t", str(threads), "-d", "reads.fq", "-r", "ref.fa", "-w", "./wrk", "-o", "o.txt", "-p", "p.txt"] + list(args)
    t0 = time.time()
    subprocess.run(cmd, cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return time.time() - t0


def main():
    ref, reads = inputs()
    d = tempfile.mkdtemp(prefix="ag2_map250_")
    synth.write_fasta(os.path.join(d, "ref.fa"), "chr1", ref)
    synth.write_fastq(os.path.join(d, "reads.fq"), [np.frombuffer(r, np.uint8) for r in reads])
    wall = run_reference(d)
    tf = open(os.path.join(d, "wrk", "1.r"), "rb").read()
    sha = lambda p: hashlib.sha256(open(os.path.join(d, p), "rb").read()).hexdigest()
    gold = {"seed": SEED, "ref_len": REF_LEN, "n_reads": N_READS, "args": "-t 1 -b 1 -l 0.5 -u 2.0 -z 200 -y 0.9",
            "reads_sha256": hashlib.sha256(b"\n".join(reads)).hexdigest(),
            "ref_sha256": hashlib.sha256(ref.tobytes()).hexdigest(),
            "thread_file_sha256": hashlib.sha256(tf).hexdigest(), "o_sha256": sha("o.txt"), "p_sha256": sha("p.txt"),
            "records": record_digest(tf), "config_txt": open(os.path.join(d, "config.txt")).read().splitlines()[5:],
            "reference_wall_s": round(wall, 1)}
    json.dump(gold, open(GOLDEN_JSON, "w"), indent=0)
    print(d, len(gold["records"]), "records", wall, "s")


if __name__ == "__main__":
    main()
