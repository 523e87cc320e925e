"""Full-path e2e with 1 vs 2 contexts on one GPU (each context: its own host thread, half of the reads):
does a second context hide the copies and the launch tails of the first?  python experiments/two_ctx.py [--reads 250000]"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=250000)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    import torch
    from aligngraph2_b200.lib import RECORD_DTYPE
    from aligngraph2_b200.mecat2ref import Mecat2RefDevice
    args = argparse.Namespace(full_ref_len=250_000_000, seed=20261017, tlen=10000)
    d = bench.full_path_inputs(args, 0, "cuda", a.reads)
    ref = d["ref"].cpu().numpy()
    h_bases = torch.empty(d["bases"].numel(), dtype=torch.uint8, pin_memory=True)
    h_bases.copy_(d["bases"])
    off = d["offsets"].cpu().numpy()
    del d
    torch.cuda.empty_cache()
    bases = h_bases.numpy()
    for nctx in (1, 2, 3):
        devs = [Mecat2RefDevice(0) for _ in range(nctx)]
        cuts = [int(x) for x in np.linspace(0, a.reads, nctx + 1)]
        parts = []
        for k, dev in enumerate(devs):
            lo, hi = cuts[k], cuts[k + 1]
            o = (off[lo:hi + 1] - off[lo]).astype(np.int64)
            b = bases[off[lo]:off[hi]]
            dev.load_reference(ref)
            dev.load_reads(bases=bases[:off[min(a.reads, 100000)]], offsets=off[:min(a.reads, 100000) + 1])   # the index comes from the first 100 k reads of the batch
            dev.build_index(200, 0.5, 2.0)
            rec = torch.empty((hi - lo + 16) * RECORD_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
            ops = torch.empty(int((off[hi] - off[lo]) * 1.1) // 16 + 4096, dtype=torch.int32, pin_memory=True)
            parts.append((dev, b, o, rec, ops))
        aligned = [0] * nctx

        def work(k):
            dev, b, o, rec, ops = parts[k]
            dev.load_reads(bases=b, offsets=o)
            n = dev.map_reads_only(10, 1)
            r = rec.numpy().view(RECORD_DTYPE)[:n]
            dev.map_fetch_packed_into(r, ops.numpy().view(np.uint32))
            aligned[k] = int((r["qe"].astype(np.int64) - r["qb"]).sum())

        def step():
            th = [threading.Thread(target=work, args=(k,)) for k in range(nctx)]
            for t in th:
                t.start()
            for t in th:
                t.join()
        step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / a.steps
        print(json.dumps({"contexts": nctx, "ms_per_step": ms, "gbp_per_s": sum(aligned) / ms / 1e6}), flush=True)
        for dev in devs:
            dev.close()


if __name__ == "__main__":
    main()
