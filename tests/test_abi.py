"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol that
include/ag2_b200.h declares.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from aligngraph2_b200 import build, lib as L
    build.build()
    return L.load()


def declared_symbols(header="ag2_b200.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ag2_[a-z0-9_]+)\s*\(", text)))


def test_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ag2_b200.h but not exported"
    from aligngraph2_b200.lib import EXPORTS
    assert sorted(EXPORTS) == names


def test_exports_every_declared_pagraph_symbol(lib):
    names = declared_symbols("ag2_pagraph.h")
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ag2_pagraph.h but not exported"
    from aligngraph2_b200.pagraph import EXPORTS, ALN_DTYPE, Params, Stats
    assert sorted(EXPORTS) == names
    assert ALN_DTYPE.itemsize == 72 and C.sizeof(Params) == 48 and C.sizeof(Stats) == 152


def test_pagraph_refuses_without_gpu(lib, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from aligngraph2_b200 import pagraph
    from aligngraph2_b200.lib import Ag2Error
    h = C.c_void_p()
    assert pagraph._L().ag2_pg_create(0, C.byref(h)) == -1   # AG2_ENODEV
    with pytest.raises(Ag2Error):
        pagraph.Job("a", "b", "c", str(tmp_path), "d")


def test_struct_layouts_match_header(lib):
    from aligngraph2_b200.lib import CANDIDATE_DTYPE, RECORD_DTYPE, ExtendStats
    assert CANDIDATE_DTYPE.itemsize == 24
    assert RECORD_DTYPE.itemsize == 56
    assert C.sizeof(ExtendStats) == 88
    from aligngraph2_b200.lib import MapStats
    assert C.sizeof(MapStats) == 7 * 8 + 9 * 8      # ag2_map_stats: 7 doubles, 9 int64


def test_version_and_no_cpu_fallback(lib):
    assert b"sm_100a" in lib.ag2_version()
    import torch
    if not torch.cuda.is_available():
        ctx = C.c_void_p()
        assert lib.ag2_ctx_create(0, C.byref(ctx)) == -1  # AG2_ENODEV: the product refuses to run without a GPU
        from aligngraph2_b200.mecat2ref import Mecat2RefDevice
        from aligngraph2_b200.lib import Ag2Error
        with pytest.raises(Ag2Error):
            Mecat2RefDevice(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "aligngraph2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower() or f in ("synth.py",), f"{f} mentions the oracle"
