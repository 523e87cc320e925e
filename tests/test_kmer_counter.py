"""PAGraph kmer_counter: oracle vs the unmodified reference binary, CUDA path vs the oracle,
drop-in executable vs the reference binary."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from aligngraph2_b200 import synth



def _reads(seed=3, n=40, tl=3000):
    d = synth.make_batch_torch(seed, 200_000, n, tl)
    bases = d["bases"].numpy().copy()
    off = d["offsets"].numpy()
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, len(bases), size=200)
    bases[idx[:100]] = ord("N")          # anything but acgt counts as A
    bases[idx[100:]] |= 0x20             # lower case counts like upper case
    # a homopolymer read and a read shorter than k
    extra = [np.frombuffer(b"A" * 500 + b"ACGT" * 50, dtype=np.uint8), np.frombuffer(b"ACGTAC", dtype=np.uint8)]
    reads = [bases[off[i]:off[i + 1]] for i in range(n)] + extra
    offs = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in reads], out=offs[1:])
    return np.concatenate(reads), offs


def _write_fastq(path, bases, offs):
    with open(path, "wb") as f:
        for i in range(len(offs) - 1):
            rd = bases[offs[i]:offs[i + 1]].tobytes()
            f.write(b"@r%d\n" % i + rd + b"\n+\n" + b"I" * len(rd) + b"\n")


@pytest.mark.parametrize("k,threshold", [(11, 0.2), (9, 0.05), (12, 0.5)])
def test_oracle_matches_reference_binary(tmp_path, k, threshold):
    from oracle import binding
    if not os.path.exists(binding.REF_KMER_COUNTER):
        pytest.skip("oracle/_ref/kmer_counter not built on this box")
    binding.build(ref=False)
    bases, offs = _reads()
    _write_fastq(tmp_path / "r.fq", bases, offs)
    subprocess.run([binding.REF_KMER_COUNTER, "-t", "1", "-i", "r.fq", "-o", "ref.bin", "-k", str(k), "-m", str(threshold)],
                   cwd=tmp_path, check=True)
    ref = np.fromfile(tmp_path / "ref.bin", dtype=np.uint64)
    codes, cut = binding.solid_kmers(bases.tobytes(), offs, k, threshold)
    assert ref[0] == k and np.array_equal(ref[1:], codes)


def test_oracle_golden_digest():
    # digest of the reference binary's output for _reads(), k = 11 (tests/golden/kmer_counter_k11.sha256)
    from oracle import binding
    binding.build(ref=False)
    bases, offs = _reads()
    codes, cut = binding.solid_kmers(bases.tobytes(), offs, 11, 0.2)
    blob = np.uint64(11).tobytes() + codes.tobytes()
    want = open(os.path.join(os.path.dirname(__file__), "golden", "kmer_counter_k11.sha256")).read().split()[0]
    assert hashlib.sha256(blob).hexdigest() == want


@pytest.mark.gpu
@pytest.mark.parametrize("k,threshold", [(11, 0.2), (9, 0.05), (13, 0.2), (14, 0.2)])
def test_cuda_matches_oracle(k, threshold):
    from aligngraph2_b200.mecat2ref import Mecat2RefDevice
    from oracle import binding
    bases, offs = _reads()
    dev = Mecat2RefDevice(0)
    dev.load_reads(bases=bases, offsets=offs)
    got, cut = dev.solid_kmers(k, threshold)
    exp, ecut = binding.solid_kmers(bases.tobytes(), offs, k, threshold)
    dev.close()
    assert cut == ecut
    assert np.array_equal(got, exp)


@pytest.mark.gpu
def test_dropin_binary_matches_reference_binary(tmp_path):
    from aligngraph2_b200 import build
    from oracle import binding
    if not os.path.exists(binding.REF_KMER_COUNTER):
        pytest.skip("oracle/_ref/kmer_counter not built on this box")
    exe = build.build_host(name="kmer_counter")
    bases, offs = _reads(seed=8, n=300, tl=5000)
    _write_fastq(tmp_path / "r.fq", bases, offs)
    subprocess.run([binding.REF_KMER_COUNTER, "-t", "1", "-i", "r.fq", "-o", "ref.bin", "-k", "12"], cwd=tmp_path, check=True)
    r = subprocess.run([exe, "-t", "8", "--in", "r.fq", "-o", "gpu.bin", "-k12"], cwd=tmp_path, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    assert (tmp_path / "ref.bin").read_bytes() == (tmp_path / "gpu.bin").read_bytes()


@pytest.mark.gpu
@pytest.mark.parametrize("devices", ["0,0,0", None])
def test_dropin_binary_sharded_over_devices(tmp_path, devices):
    """the read batches go round robin over several contexts (AG2_DEVICES names GPU 0 three times on a one-GPU
    box; None = every visible GPU), each counts into its own 4^k table, ag2_kmer_merge sums them over peer memory before the
    cut.  The file must be the reference binary's."""
    from aligngraph2_b200 import build
    from oracle import binding
    if not os.path.exists(binding.REF_KMER_COUNTER):
        pytest.skip("oracle/_ref/kmer_counter not built on this box")
    exe = build.build_host(name="kmer_counter")
    bases, offs = _reads(seed=9, n=400, tl=4000)
    _write_fastq(tmp_path / "r.fq", bases, offs)
    subprocess.run([binding.REF_KMER_COUNTER, "-t", "1", "-i", "r.fq", "-o", "ref.bin", "-k", "12"], cwd=tmp_path, check=True)
    env = {k: v for k, v in os.environ.items() if k != "AG2_DEVICES"}
    env["AG2_KMER_BATCH_BYTES"] = "100000"      # ~ 25 reads per batch: every context gets several
    if devices:
        env["AG2_DEVICES"] = devices
    r = subprocess.run([exe, "-t", "8", "-i", "r.fq", "-o", "gpu.bin", "-k", "12"], cwd=tmp_path, env=env, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    assert (tmp_path / "ref.bin").read_bytes() == (tmp_path / "gpu.bin").read_bytes()
