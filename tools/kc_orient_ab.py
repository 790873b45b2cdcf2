"""Degree- vs degeneracy-oriented DAG for the clique kernels (BASELINE.json configs[2] names the degeneracy ordering;
the reference pipeline is getDegeneracyOrderingDanischHeap -> InduceDirectedGraph -> EP_kclisting,
gms/algorithms/non_set_based/k_clique_list/bench_helper.h:33-38).

    python tools/kc_orient_ab.py [--scale 22] [--ks 5,6]

One JSON line per (orientation, k): count, seconds, max d+, class histogram.  GMSB_KCLIQUE_TRACE=1 adds the per-class
device times and clique counts on stderr."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gms_b200 as G  # noqa: E402


def hist(dag):
    off, _ = dag.export_csr()
    d = np.diff(off)
    return {"max_dplus": int(d.max()), "gt32": int((d > 32).sum()), "gt64": int((d > 64).sum()),
            "gt128": int((d > 128).sum()), "gt256": int((d > 256).sum()), "gt512": int((d > 512).sum()),
            "sum_d2": float((d.astype(np.float64) ** 2).sum())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--ks", default="5,6")
    args = ap.parse_args()
    G.set_device(0)
    src, dst = G.generate_rmat(args.scale)
    g = G.Graph.from_edgelist(src, dst, True)
    del src, dst
    n = g.n
    t0 = time.time()
    rank_deg = g.degree_order(rank_format=True)
    t_deg = time.time() - t0
    t0 = time.time()
    rank_dgn = g.degeneracy_rank()                      # reference convention: first removed = highest rank
    t_dgn = time.time() - t0
    # low -> high orientation with d+ <= core number: first removed gets rank 0
    dags = {"degree": (g.orient(rank_deg), t_deg), "degeneracy": (g.orient((n - 1 - rank_dgn).astype(np.int32)), t_dgn)}
    for name, (dag, t_order) in dags.items():
        h = hist(dag)
        dag.kclique_count(3)
        for k in [int(x) for x in args.ks.split(",")]:
            G.synchronize()
            t0 = time.time()
            c = dag.kclique_count(k)
            sec = time.time() - t0
            print(json.dumps({"orientation": name, "scale": args.scale, "k": k, "count": c, "seconds": round(sec, 4),
                              "order_seconds": round(t_order, 4), **h}), flush=True)


if __name__ == "__main__":
    main()
