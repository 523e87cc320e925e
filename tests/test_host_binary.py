"""The drop-in `mecat2ref` executable (aligngraph2_b200/host/mecat2ref_main.cpp).

CPU: with AG2_SKIP_MAP=1 the GPU stage is skipped and the program runs its file handling (0.fq, ref.fq,
chrindex.txt, config.txt) plus result_combine / polish_result on the reference's own thread file; every file must be
byte-identical to what the unmodified reference binary wrote (tests/golden/mapper_stress_ref.*).
GPU: the whole program, same comparison, thread file included."""
import hashlib
import importlib.util
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, golden_ref_outputs

ARGS = ["-d", "reads.fq", "-r", "ref.fa", "-b", "1", "-w", "./wrk", "-o", "o.txt", "-p", "p.txt", "-l", "0.5", "-u", "2.0",
        "-z", "200", "-y", "0.9"]


@pytest.fixture(scope="module")
def binary():
    from aligngraph2_b200 import build
    build.build()
    return build.build_host()


def _inputs(d):
    spec = importlib.util.spec_from_file_location("gen_mapper_golden", os.path.join(GOLDEN, "gen_mapper_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    chroms, reads = gen.build()
    z = np.load(os.path.join(GOLDEN, "mapper_stress.npz"))
    assert b"".join(reads) == z["bases"].tobytes()       # the generator still makes the committed fixture
    gen.write_inputs(str(d), chroms, reads)


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def _check_outputs(d, gold, thread_file: bool):
    hashes = json.load(open(os.path.join(GOLDEN, "mapper_stress_ref.json")))
    assert _sha(d / "wrk" / "0.fq") == hashes["wrk/0.fq"]
    assert _sha(d / "wrk" / "ref.fq") == hashes["wrk/ref.fq"]
    assert (d / "wrk" / "chrindex.txt").read_bytes() == gold["wrk/chrindex.txt"]
    if thread_file:
        assert (d / "wrk" / "1.r").read_bytes() == gold["wrk/1.r"]
    assert (d / "o.txt").read_bytes() == gold["o.txt"]
    assert (d / "p.txt").read_bytes() == gold["p.txt"]
    cfg = (d / "config.txt").read_text().splitlines()
    assert cfg[:7] == ["./wrk", "ref.fa", "reads.fq", "o.txt", "p.txt", "1\t90", "2"]
    assert cfg[7].startswith("The Building read Index Time:") and cfg[9].startswith("The Mapping Time:")
    assert (d / "p.txt.config").exists()


def test_files_combine_and_polish_match_reference(binary, tmp_path):
    gold = golden_ref_outputs()
    _inputs(tmp_path)
    (tmp_path / "wrk").mkdir()
    (tmp_path / "wrk" / "1.r").write_bytes(gold["wrk/1.r"])
    r = subprocess.run([binary] + ARGS, cwd=tmp_path, env=dict(os.environ, AG2_SKIP_MAP="1"), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    _check_outputs(tmp_path, gold, thread_file=False)


def test_usage_errors_exit_1(binary, tmp_path):
    r = subprocess.run([binary, "-d", "x.fq"], cwd=tmp_path, capture_output=True)
    assert r.returncode == 1 and b"reference must be specified" in r.stderr
    r = subprocess.run([binary] + ARGS + ["-n", "0"], cwd=tmp_path, capture_output=True)
    assert r.returncode == 1 and b"candidates must be > 0" in r.stderr


def test_without_gpu_exits_1(binary, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _inputs(tmp_path)
    r = subprocess.run([binary] + ARGS, cwd=tmp_path, capture_output=True)
    assert r.returncode == 1 and b"no CPU path" in r.stderr     # AlignGraph2.py:280-296 then falls back to vanilla mecat2ref


@pytest.mark.gpu
def test_whole_program_matches_reference(binary, tmp_path):
    gold = golden_ref_outputs()
    _inputs(tmp_path)
    r = subprocess.run([binary] + ARGS + ["-t", "4"], cwd=tmp_path, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    cfg = (tmp_path / "config.txt").read_text().splitlines()
    assert cfg[5] == "4\t90"
    (tmp_path / "config.txt").write_text("\n".join(cfg[:5] + ["1\t90"] + cfg[6:]) + "\n")
    _check_outputs(tmp_path, gold, thread_file=True)
    assert (tmp_path / "wrk" / "4.r").read_bytes() == b""


@pytest.mark.gpu
@pytest.mark.parametrize("devices", ["0,0", "0,0,0,0,0"])
def test_reads_sharded_over_devices_give_the_same_files(binary, tmp_path, devices):
    """every load_fastq batch is cut into contiguous read ranges, one per device (its own context, index copy
    and host thread); the thread file is written in device order.  AG2_DEVICES may name a device more than once, so the
    sharding runs on a one-GPU box: all files must equal the reference's, whatever the number of shards."""
    gold = golden_ref_outputs()
    _inputs(tmp_path)
    r = subprocess.run([binary] + ARGS, cwd=tmp_path, env=dict(os.environ, AG2_DEVICES=devices), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    _check_outputs(tmp_path, gold, thread_file=True)


@pytest.mark.gpu
def test_whole_program_matches_reference_binary_run(binary, tmp_path):
    from oracle import binding
    if not os.path.exists(binding.REF_BIN):
        pytest.skip("oracle/_ref/mecat2ref not built on this box")
    spec = importlib.util.spec_from_file_location("gen_mapper_golden", os.path.join(GOLDEN, "gen_mapper_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    chroms, reads = gen.build(seed=99)
    a, b = tmp_path / "ref", tmp_path / "gpu"
    for d in (a, b):
        d.mkdir()
        gen.write_inputs(str(d), chroms, reads[:60])
    args = [x if x != "1" else "2" for x in ARGS]       # -b 2
    subprocess.run([binding.REF_BIN, "-t", "1"] + args, cwd=a, check=True, capture_output=True)
    r = subprocess.run([binary] + args, cwd=b, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    for name in ("wrk/0.fq", "wrk/ref.fq", "wrk/chrindex.txt", "wrk/1.r", "o.txt", "p.txt"):
        assert (a / name).read_bytes() == (b / name).read_bytes(), name


@pytest.mark.gpu
def test_more_than_one_load_fastq_batch(binary, tmp_path):
    # 100 200 short reads: load_fastq hands out 100 001 reads, then the rest (impl_large.cpp:1965-1991), and the read
    # index only sees the first 100 000 (:277).  Whole program against a fresh `-t 1` run of the reference binary.
    from oracle import binding
    if not os.path.exists(binding.REF_BIN):
        pytest.skip("oracle/_ref/mecat2ref not built on this box")
    from aligngraph2_b200 import synth
    d = synth.make_batch_torch(4711, 1_500_000, 100_200, 1150, device="cuda")
    ref, bases, off = d["ref"].cpu().numpy(), d["bases"].cpu().numpy(), d["offsets"].cpu().numpy()
    a, b = tmp_path / "ref", tmp_path / "gpu"
    for dd in (a, b):
        dd.mkdir()
        synth.write_fasta(str(dd / "ref.fa"), "chr1", ref)
        with open(dd / "reads.fq", "wb") as f:
            for i in range(len(off) - 1):
                rd = bases[off[i]:off[i + 1]].tobytes()
                f.write(b"@r%d\n" % i + rd + b"\n+\n" + b"I" * len(rd) + b"\n")
    subprocess.run([binding.REF_BIN, "-t", "1"] + ARGS, cwd=a, check=True, capture_output=True)
    r = subprocess.run([binary] + ARGS, cwd=b, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    for name in ("wrk/0.fq", "wrk/chrindex.txt", "wrk/1.r", "o.txt", "p.txt"):
        assert (a / name).read_bytes() == (b / name).read_bytes(), name
    assert (b / "p.txt").read_bytes().count(b"\n") > 3 * 99_000
    # the same with both batches sharded over three contexts (the index still comes from the whole first batch)
    c = tmp_path / "gpu3"
    c.mkdir()
    for name in ("ref.fa", "reads.fq"):
        os.link(b / name, c / name)
    r = subprocess.run([binary] + ARGS, cwd=c, env=dict(os.environ, AG2_DEVICES="0,0,0"), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    for name in ("wrk/1.r", "o.txt", "p.txt"):
        assert (b / name).read_bytes() == (c / name).read_bytes(), name


def _records(path):
    """multiset of 3-line records of a ref-format / thread file"""
    lines = open(path, "rb").read().split(b"\n")
    return sorted(b"\n".join(lines[k:k + 3]) for k in range(0, len(lines) - 2, 3))


@pytest.mark.gpu
def test_baseline_config0_against_reference_t1_and_t8(binary, tmp_path):
    """BASELINE configs[0] exactly: 1 000 synthetic CLR reads (10 kb templates, 15 % error) vs a 1 Mb reference, seed
    20261017, `-z 200`.  The drop-in executable's `-p` / `-o` files, as multisets of records, must equal what the
    UNMODIFIED reference binary writes with `-t 1` AND with `-t 8` (per-read results do not depend on -t, only
    the order of the records does), and byte for byte what `-t 1` writes; running ours with `-t 8` changes nothing."""
    from oracle import binding
    if not os.path.exists(binding.REF_BIN):
        pytest.skip("oracle/_ref/mecat2ref not built on this box")
    from aligngraph2_b200 import synth
    ref, reads, _ = synth.make_dataset(20261017, 1_000_000, 1000, 10000)
    dirs = {k: tmp_path / k for k in ("ref1", "ref8", "gpu1", "gpu8")}
    for d in dirs.values():
        d.mkdir()
        synth.write_fasta(str(d / "ref.fa"), "chr1", ref)
        synth.write_fastq(str(d / "reads.fq"), reads)
    subprocess.run([binding.REF_BIN, "-t", "1"] + ARGS, cwd=dirs["ref1"], check=True, capture_output=True)
    subprocess.run([binding.REF_BIN, "-t", "8"] + ARGS, cwd=dirs["ref8"], check=True, capture_output=True)
    for k, t in (("gpu1", "1"), ("gpu8", "8")):
        r = subprocess.run([binary, "-t", t] + ARGS, cwd=dirs[k], capture_output=True)
        assert r.returncode == 0, r.stderr.decode()
    want_p, want_o = _records(dirs["ref1"] / "p.txt"), _records(dirs["ref1"] / "o.txt")
    assert len(want_p) >= 990
    assert _records(dirs["ref8"] / "p.txt") == want_p and _records(dirs["ref8"] / "o.txt") == want_o     # F6 holds on this input
    for k in ("gpu1", "gpu8"):
        assert _records(dirs[k] / "p.txt") == want_p, k
        assert _records(dirs[k] / "o.txt") == want_o, k
        for name in ("wrk/0.fq", "wrk/ref.fq", "wrk/chrindex.txt", "wrk/1.r", "o.txt", "p.txt"):
            assert (dirs["ref1"] / name).read_bytes() == (dirs[k] / name).read_bytes(), (k, name)


def test_fastq_token_semantics_match_reference(binary, tmp_path):
    """The reference reads FASTQ with fscanf("%[^\\n]s") / fscanf("%s\\n") pairs (mecat2ref.cpp:317): sequence and quality are
    whitespace-delimited TOKENS and all whitespace behind them, blank lines included, is swallowed.  <wrk>/0.fq (ids and
    lengths) of the drop-in executable must equal the reference binary's on a file with CRLF line ends, trailing blanks,
    blank lines between records, a leading blank before a sequence, and no newline at the end."""
    from oracle import binding
    if not os.path.exists(binding.REF_BIN):
        pytest.skip("oracle/_ref/mecat2ref not built on this box")
    rng = np.random.default_rng(11)
    seq = lambda n: bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=n))
    recs = [b"@r0 first\n" + seq(300) + b"\n+\n" + b"I" * 300 + b"\n",                       # clean
            b"@r1\r\n" + seq(250) + b"\r\n+\r\n" + b"I" * 250 + b"\r\n",                      # CRLF
            b"@r2\n" + seq(200) + b"  \t\n+r2\n" + b"I" * 200 + b" \n",                       # trailing blanks
            b"\n\n@r3\n" + seq(180) + b"\n+\n" + b"I" * 180 + b"\n\n\n",                      # blank lines around the record
            b"@r4\n  " + seq(150) + b"\n+\n" + b"I" * 150 + b"\n",                            # leading blanks before the sequence
            b"@r5\n" + seq(120) + b"\n\n+\n" + b"I" * 120 + b"\n",                            # blank line inside the record
            b"@r6\n" + seq(100) + b"\n+\n" + b"I" * 100]                                      # no newline at the end of the file
    ref = b">chr1\n" + seq(5000) + b"\n"
    outs = {}
    for who, exe, env in (("ref", binding.REF_BIN, os.environ), ("ours", binary, dict(os.environ, AG2_SKIP_MAP="1"))):
        d = tmp_path / who
        d.mkdir()
        (d / "ref.fa").write_bytes(ref)
        (d / "reads.fq").write_bytes(b"".join(recs))
        if who == "ours":
            (d / "wrk").mkdir()
            (d / "wrk" / "1.r").write_bytes(b"")
        r = subprocess.run([exe, "-t", "1"] + ARGS, cwd=d, env=env, capture_output=True)
        assert r.returncode == 0, r.stderr.decode()
        outs[who] = (d / "wrk" / "0.fq").read_bytes()
    assert outs["ours"] == outs["ref"]
    assert outs["ref"].count(b"\n") == 7


@pytest.mark.parametrize("fmt", ["1", "2"])
def test_m4_and_sam_formats_match_reference(binary, tmp_path, fmt):
    """`-m 1` (m4) and `-m 2` (SAM) of result_combine / polish_result (output.cpp:45-186): the reference binary maps the stress
    fixture and writes -o / -p in that format; the drop-in executable makes the same files from the reference's thread file
    (AG2_SKIP_MAP: no GPU needed -- the GPU run goes through the same formatters, see the test below).  SAM's @PG line
    carries argv[0], which differs by construction."""
    from oracle import binding
    if not os.path.exists(binding.REF_BIN):
        pytest.skip("oracle/_ref/mecat2ref not built on this box")
    args = ARGS + ["-m", fmt]
    a, b = tmp_path / "ref", tmp_path / "ours"
    for d in (a, b):
        d.mkdir()
        _inputs(d)
    subprocess.run([binding.REF_BIN, "-t", "1"] + args, cwd=a, check=True, capture_output=True)
    (b / "wrk").mkdir()
    (b / "wrk" / "1.r").write_bytes((a / "wrk" / "1.r").read_bytes())
    r = subprocess.run([binary] + args, cwd=b, env=dict(os.environ, AG2_SKIP_MAP="1"), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    norm = lambda blob: b"\n".join(ln for ln in blob.split(b"\n") if not ln.startswith(b"@PG"))
    for name in ("o.txt", "p.txt"):
        want, got = (a / name).read_bytes(), (b / name).read_bytes()
        assert len(want) > 1000
        assert norm(got) == norm(want), name


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["1", "2"])
def test_m4_and_sam_formats_whole_program(binary, tmp_path, fmt):
    """The same formats from a mapping run (records formatted from memory by the ResultWriter)."""
    from oracle import binding
    if not os.path.exists(binding.REF_BIN):
        pytest.skip("oracle/_ref/mecat2ref not built on this box")
    args = ARGS + ["-m", fmt]
    a, b = tmp_path / "ref", tmp_path / "gpu"
    for d in (a, b):
        d.mkdir()
        _inputs(d)
    subprocess.run([binding.REF_BIN, "-t", "1"] + args, cwd=a, check=True, capture_output=True)
    r = subprocess.run([binary] + args, cwd=b, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    norm = lambda blob: b"\n".join(ln for ln in blob.split(b"\n") if not ln.startswith(b"@PG"))
    for name in ("o.txt", "p.txt"):
        assert norm((b / name).read_bytes()) == norm((a / name).read_bytes()), name
