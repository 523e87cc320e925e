"""The product's per-read pipeline logic around the extensions -- seed_device.cuh (seeding, candidates) and
rescue_device.cuh (plan_read / finish_read: rescue planning, linking, output choice, second pass), compiled unchanged with
-DAG2_EMU by tests/emu/build.sh -- on the CPU: the thread file it writes for the stress fixture must be byte-identical to
the `<wrk>/1.r` the unmodified reference binary wrote (tests/golden/mapper_stress_ref.tar.xz).  The extensions themselves
are done by the oracle here (the kernels are checked in tests/test_emu_xdrop.py and on the GPU)."""
import ctypes as C
import os
import subprocess

import numpy as np

from conftest import GOLDEN, golden_ref_outputs

EMU_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")


def test_emulated_per_read_path_writes_the_reference_thread_file(tmp_path):
    from oracle.binding import IndexOracle
    subprocess.run([os.path.join(EMU_DIR, "build.sh")], check=True)
    E = C.CDLL(os.path.join(EMU_DIR, "libemu_map.so"))
    E.emu_map_batch.restype = C.c_long
    E.emu_map_batch.argtypes = [C.c_char_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_void_p,
                                C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_void_p]
    z = np.load(os.path.join(GOLDEN, "mapper_stress.npz"))
    bases, offs = z["bases"].tobytes(), z["offsets"].astype(np.int64)
    io = IndexOracle(z["genome"].tobytes(), bases, offs)
    n = len(offs) - 1
    ids = np.arange(1, n + 1, dtype=np.int32)
    stats = np.zeros(2, dtype=np.int64)
    out = str(tmp_path / "emu.r")
    written = E.emu_map_batch(io.ref, len(io.ref), io.cnt.ctypes.data, io.off.ctypes.data, io.pos.ctypes.data, io.vote.ctypes.data, 200,
                              bases, offs.ctypes.data, ids.ctypes.data, n, 10, 1, out.encode(), stats.ctypes.data)
    golden = golden_ref_outputs()["wrk/1.r"]
    assert open(out, "rb").read() == golden
    assert written == golden.count(b"\n") // 3
    assert stats[0] >= 1 and stats[1] >= 1          # rescue extensions and second-pass reads are exercised
