"""The drop-in `mecat2ref` executable's HOST flow in the GPU-less container: load_fastq batching, the cut of a batch into
per-device read ranges (one host thread and one context per device), read ids, record order, thread file, result_combine and
polish_result -- with a TEST DOUBLE of the C ABI behind it (tests/host/fake_ag2_lib.cpp: the oracle answers the host's calls)
instead of the CUDA library.  Every file must equal what the unmodified reference binary wrote, whatever the device count.
The real library on a real GPU runs the same comparison in tests/test_host_binary.py (`-m gpu`)."""
import os
import shutil
import subprocess

import pytest

from conftest import golden_ref_outputs
from test_host_binary import ARGS, _check_outputs, _inputs

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def staged(tmp_path_factory):
    """<tmp>/libag2_b200.so = the test double, <tmp>/bin/mecat2ref = the product's executable source linked against it (rpath
    $ORIGIN/..).  Needs g++ only: neither nvcc nor the CUDA library."""
    from oracle import binding
    binding.build(ref=False)
    d = tmp_path_factory.mktemp("fake_lib")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if not cxx:
        pytest.skip("no C++ compiler on this box")
    oracle_dir = os.path.join(ROOT, "oracle")
    subprocess.run([cxx, "-O1", "-std=c++17", "-Wall", "-fPIC", "-shared", "-pthread", "-o", str(d / "libag2_b200.so"),
                    os.path.join(HERE, "host", "fake_ag2_lib.cpp"), "-L" + oracle_dir, "-lag2_oracle", "-Wl,-rpath," + oracle_dir], check=True)
    (d / "bin").mkdir()
    subprocess.run([cxx, "-O2", "-std=c++17", "-Wall", "-pthread", "-o", str(d / "bin" / "mecat2ref"),
                    os.path.join(ROOT, "aligngraph2_b200", "host", "mecat2ref_main.cpp"), "-L" + str(d), "-lag2_b200",
                    "-Wl,-rpath,$ORIGIN/.."], check=True)
    return str(d / "bin" / "mecat2ref")


@pytest.mark.parametrize("devices", [1, 3, 7])
def test_host_flow_with_test_double(staged, tmp_path, devices):
    gold = golden_ref_outputs()
    _inputs(tmp_path)
    env = {k: v for k, v in os.environ.items() if k not in ("AG2_DEVICES", "AG2_SKIP_MAP", "LD_LIBRARY_PATH")}
    r = subprocess.run([staged] + ARGS, cwd=tmp_path, env=dict(env, FAKE_AG2_DEVICES=str(devices)), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    _check_outputs(tmp_path, gold, thread_file=True)


@pytest.mark.skipif(not os.environ.get("AG2_SLOW_TESTS"), reason="13 minutes of single-threaded oracle; set AG2_SLOW_TESTS=1 "
                    "(the GPU suite runs the same input in test_host_binary.py::test_more_than_one_load_fastq_batch)")
def test_host_flow_across_load_fastq_batches(staged, tmp_path):
    """100 200 short reads: load_fastq hands out 100 001 reads, then the rest (impl_large.cpp:1965-1991), the read index only
    sees the first 100 000 (:277); the executable parses the second batch on a second thread while the first is mapped.
    Files against a fresh run of the unmodified reference binary; two "devices", so both batches are cut in two."""
    import numpy as np
    from aligngraph2_b200 import synth
    from oracle import binding
    if not os.path.exists(binding.REF_BIN):
        if not os.path.isdir("/root/reference/mecat_plus"):
            pytest.skip("oracle/_ref/mecat2ref not built and no /root/reference here")
        binding.build(ref=True)
    d = synth.make_batch_torch(4711, 1_500_000, 100_200, 1150)
    ref, bases, off = d["ref"].numpy(), d["bases"].numpy(), d["offsets"].numpy()
    a, b = tmp_path / "ref", tmp_path / "fake"
    a.mkdir()
    synth.write_fasta(str(a / "ref.fa"), "chr1", ref)
    with open(a / "reads.fq", "wb") as f:
        for i in range(len(off) - 1):
            rd = bases[off[i]:off[i + 1]].tobytes()
            f.write(b"@r%d\n" % i + rd + b"\n+\n" + b"I" * len(rd) + b"\n")
    b.mkdir()
    for name in ("ref.fa", "reads.fq"):
        os.link(a / name, b / name)
    subprocess.run([binding.REF_BIN, "-t", "1"] + ARGS, cwd=a, check=True, capture_output=True)
    env = {k: v for k, v in os.environ.items() if k not in ("AG2_DEVICES", "AG2_SKIP_MAP", "LD_LIBRARY_PATH")}
    r = subprocess.run([staged] + ARGS, cwd=b, env=dict(env, FAKE_AG2_DEVICES="2"), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    for name in ("wrk/0.fq", "wrk/chrindex.txt", "wrk/1.r", "o.txt", "p.txt"):
        assert (a / name).read_bytes() == (b / name).read_bytes(), name
    assert (b / "p.txt").read_bytes().count(b"\n") > 3 * 99_000
