"""torchrun entry of the multi-GPU A-Bruijn build: one process per GPU, reads sharded over the ranks.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      -m aligngraph2_b200.pagraph_dist_main -k solid.bin -c ctg.fasta -R ref.fasta -p <pre dir> -a c2r.ref -o graph.txt

Every rank builds the vertices it owns (aligngraph2_b200.pagraph.build_distributed); the tables are merged with one
all-gather (gather_graph) and rank 0 writes the dump of all config blocks to -o.
"""
from __future__ import annotations

import argparse
import os

import torch
import torch.distributed as dist

from . import pagraph


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("-k", required=True)
    ap.add_argument("-c", required=True)
    ap.add_argument("-R", required=True)
    ap.add_argument("-p", required=True)
    ap.add_argument("-a", required=True)
    ap.add_argument("-o", required=True)
    ap.add_argument("--epsilon", type=int, default=10)
    ap.add_argument("-v", type=int, default=1)
    a = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank()
    job = pagraph.Job(a.k, a.c, a.R, a.p, a.a, device=local)
    params = pagraph.default_params(a.epsilon, a.v)
    codes = job.codes()
    blob = b""
    for b in range(job.n_blocks):
        st = pagraph.build_distributed(job, b, params)
        g = pagraph.gather_graph(job)
        if rank == 0:
            blob += pagraph.graph_dump_text(g, codes, b, job.block_ref(b))
        print(f"rank {rank} block {b}: tuples {list(st.tuples)} positions {st.positions} edges {st.edges}", flush=True)
    if rank == 0:
        with open(a.o, "wb") as f:
            f.write(blob)
    job.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
