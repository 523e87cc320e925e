import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_npz_rows(name):
    z = np.load(os.path.join(GOLDEN, name))
    n = int(z["n"])
    keys = sorted({k.rsplit("_", 1)[0] for k in z.files if k.rsplit("_", 1)[-1].isdigit()})
    rows = [{k: z[f"{k}_{i}"] for k in keys} for i in range(n)]
    return z, rows


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.build(ref=False)
    return binding.Oracle()


@pytest.fixture(scope="session")
def reflib():
    from oracle import binding
    if not binding.have_ref():
        if os.path.isdir("/root/reference/mecat_plus"):
            binding.build(ref=True)
        else:
            pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    return binding.RefLib()


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


def random_block(rng):
    M = int(rng.integers(1, 719))
    rate = float(rng.choice([0.0, 0.05, 0.15, 0.3, 0.5]))
    A = rng.integers(0, 4, size=M).astype(np.uint8)
    if rng.random() < 0.15:
        A = np.tile(rng.integers(0, 4, size=int(rng.integers(1, 6))), M)[:M].astype(np.uint8)
    B = mutate(A, rate, rng)
    if rng.random() < 0.1 or len(B) == 0:
        B = rng.integers(0, 4, size=int(rng.integers(1, 719))).astype(np.uint8)
    B = B[:718]
    return A, B


def golden_ref_outputs():
    """{name: bytes} of what the unmodified reference binary wrote for the stress fixture (gen_mapper_golden.py)."""
    import lzma
    import tarfile
    out = {}
    with lzma.open(os.path.join(GOLDEN, "mapper_stress_ref.tar.xz")) as xz, tarfile.open(fileobj=xz) as tar:
        for m in tar.getmembers():
            out[m.name] = tar.extractfile(m).read()
    return out
