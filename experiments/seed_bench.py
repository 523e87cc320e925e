"""Stage timing of ag2_map_reads on the full-path workload (250 Mb reference) for profiling runs:
  python experiments/seed_bench.py --reads 50000 [--ref-len 250000000] [--steps 3]
prints ag2_map_stats per step.  Under ncu: -k regex:seed_cta_kernel -c 1."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=50000)
    ap.add_argument("--ref-len", type=int, default=250_000_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--seed-only", action="store_true")
    a = ap.parse_args()
    import torch
    from aligngraph2_b200.mecat2ref import Mecat2RefDevice
    args = argparse.Namespace(full_ref_len=a.ref_len, seed=20261017, tlen=10000)
    d = bench.full_path_inputs(args, 0, "cuda", a.reads)
    ref, bases, off = d["ref"].cpu().numpy(), d["bases"].cpu().numpy(), d["offsets"].cpu().numpy()
    del d
    torch.cuda.empty_cache()
    dev = Mecat2RefDevice(0)
    dev.load_reference(ref)
    dev.load_reads(bases=bases, offsets=off)
    dev.build_index(200, 0.5, 2.0)
    for _ in range(a.steps):
        if a.seed_only:
            import time
            t0 = time.perf_counter()
            c, n = dev.seed_candidates(0, 10)
            print(json.dumps({"seed_candidates_ms_incl_d2h": (time.perf_counter() - t0) * 1e3, "ncand": int(n.sum())}))
        else:
            dev.map_reads_only(10, 1)
            print(json.dumps(dev.map_stats()))
    dev.close()


if __name__ == "__main__":
    main()
