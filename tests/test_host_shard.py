"""Host logic of the multi-GPU runs of the drop-in executables (aligngraph2_b200/host/shard_split.h): the device
list and the contiguous, base-balanced read ranges.  Pure C++, compiled and run here without a GPU; the GPU side of it is
tests/test_host_binary.py::test_reads_sharded_over_devices_give_the_same_files."""
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("probe") / "shard_split_probe")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O1", "-std=c++17", "-Wall", "-Werror", "-o", exe, os.path.join(HERE, "host", "shard_split_probe.cpp")], check=True)
    return exe


def _split(probe, parts, lens):
    out = subprocess.run([probe, "split", str(parts)] + [str(x) for x in lens], check=True, capture_output=True, text=True).stdout
    return [tuple(int(v) for v in ln.split()) for ln in out.splitlines()]


@pytest.mark.parametrize("text,exp", [("0", [0]), ("0,1,2", [0, 1, 2]), ("0,0", [0, 0]), ("3,1,", [3, 1]), ("", []), ("x", []),
                                      ("2,x,4", [2])])
def test_device_list(probe, text, exp):
    out = subprocess.run([probe, "devices", text], check=True, capture_output=True, text=True).stdout
    assert [int(v) for v in out.split()] == exp


@pytest.mark.parametrize("parts", [1, 2, 3, 5, 8])
def test_ranges_are_contiguous_cover_the_batch_and_balance_bases(probe, parts):
    rng = np.random.default_rng(parts)
    lens = rng.integers(500, 30_000, size=400)
    r = _split(probe, parts, lens)
    assert len(r) == parts and r[0][0] == 0 and r[-1][1] == len(lens)
    assert all(a[1] == b[0] for a, b in zip(r, r[1:])) and all(lo <= hi for lo, hi in r)
    share = [int(lens[lo:hi].sum()) for lo, hi in r]
    assert sum(share) == int(lens.sum())
    assert max(share) - min(share) <= 2 * int(lens.max())          # no part is more than a read or two off the mean


def test_fewer_reads_than_parts_and_empty_reads(probe):
    r = _split(probe, 5, [1000, 1000])
    assert r[0][0] == 0 and r[-1][1] == 2 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    assert sorted(hi - lo for lo, hi in r) == [0, 0, 0, 1, 1]
    r = _split(probe, 2, [0, 0, 10, 0])                              # empty reads travel with a neighbour, none is lost
    assert r[0][0] == 0 and r[-1][1] == 4 and r[0][1] == r[1][0]
    r = _split(probe, 3, [7])
    assert r[-1][1] == 1 and sum(hi - lo for lo, hi in r) == 1
