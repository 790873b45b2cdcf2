"""k-clique counts/s on N GPUs (BASELINE.json configs[2]): one process per GPU, graph replicated, per-vertex
sub-problems dealt out round-robin, one all-reduce of the uint64 counts.

    python tools/bench_kclique.py [--scale 22] [--k 4,5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_kclique.py ...

Prints one JSON line per k on rank 0 (device time = max over ranks of the wall time around the C-ABI call)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import gms_b200 as G
    from gms_b200 import dist as gd
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--k", default="4,5")
    args = ap.parse_args()
    rank, world, local = gd.init()
    torch.cuda.set_device(local)
    G.set_device(local)
    dev = torch.device("cuda", local)
    src, dst = G.generate_rmat(args.scale)
    g = G.Graph.from_edgelist(src, dst, True)
    m = g.slots // 2
    g.kclique_count(3, rank, world)                    # builds the oriented DAG, warms the allocator
    for k in [int(x) for x in args.k.split(",")]:
        gd.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        part = g.kclique_count(k, rank, world)
        torch.cuda.synchronize()
        sec = gd.allreduce_max(time.perf_counter() - t0, device=dev)
        total, = gd.allreduce_counts([part], device=dev)
        if rank == 0:
            print(json.dumps({"metric": "kclique_counts_per_sec", "k": k, "scale": args.scale, "n_gpus": world,
                              "count": total, "seconds": sec, "value": total / sec, "unit": "cliques/s",
                              "edges_per_s": m / sec}), flush=True)
    gd.barrier()
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
