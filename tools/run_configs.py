"""Run the BASELINE.json configurations other than the headline one and print one JSON line each:

  cfg3  k-clique counting k=4..8 on R-MAT scale-22 ef16 (degree orientation, bit-matrix kernels)
  cfg4  Jaccard / common-neighbour scores over all edges of the same graph
  cfg5  triangle counting on the skewed R-MAT scale-26 (a=0.65, b=c=0.15)

    python tools/run_configs.py [--cfg 3,4,5] [--scale22 22] [--scale26 26] [--kmax 8] [--kcap-seconds 120]

Counts are printed so that they can be compared with the reference where it can be run; wall times are host
clocks around the C-ABI call (result back on the host), after one warm-up call where that is affordable.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gms_b200 as G  # noqa: E402


def emit(**kw):
    print(json.dumps(kw), flush=True)


def build(scale, a=0.57, bc=0.19):
    t0 = time.time()
    src, dst = G.generate_rmat(scale, a=a, b=bc, c=bc)
    t1 = time.time()
    g = G.Graph.from_edgelist(src, dst, True)
    G.synchronize()
    t2 = time.time()
    return g, {"scale": scale, "a": a, "n": g.n, "m": g.slots // 2, "generate_s": round(t1 - t0, 2),
               "build_gpu_s": round(t2 - t1, 3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="3,4,5")
    ap.add_argument("--scale22", type=int, default=22)
    ap.add_argument("--scale26", type=int, default=26)
    ap.add_argument("--kmax", type=int, default=8)
    ap.add_argument("--kcap-seconds", type=float, default=120.0)
    args = ap.parse_args()
    cfgs = set(int(c) for c in args.cfg.split(","))

    if cfgs & {3, 4}:
        g, info = build(args.scale22)
        emit(config="graph", **info)
        tri, st = g.tc_total_ex()
        emit(config="tc", scale=args.scale22, triangles=tri, count_ms=st["ms_count"], orient_ms=st["ms_orient"],
             max_dplus=st["max_dplus"], algorithmic_bytes=st["algorithmic_bytes"])
        if 3 in cfgs:
            prev = None
            for k in range(3, args.kmax + 1):
                if prev is not None and prev * 50 > args.kcap_seconds:
                    emit(config="cfg3-kclique", scale=args.scale22, k=k, skipped=f"predicted > {args.kcap_seconds}s "
                         f"(previous k took {prev:.1f}s; the search grows ~50x per k at scale 22, see DESIGN.md section 3)")
                    prev = prev * 50
                    continue
                t0 = time.time()
                c = g.kclique_count(k)
                dt = time.time() - t0
                prev = dt
                emit(config="cfg3-kclique", scale=args.scale22, k=k, count=c, seconds=round(dt, 4),
                     cliques_per_s=c / dt, edges_per_s=info["m"] / dt)
        if 4 in cfgs:
            for metric in ("comm_neigh", "jaccard"):
                g.edge_similarity(metric)          # warm-up (allocator, schedule)
                t0 = time.time()
                s = g.edge_similarity(metric)
                dt = time.time() - t0
                emit(config="cfg4-edge-similarity", scale=args.scale22, metric=metric, edges=len(s), seconds=round(dt, 4),
                     edges_per_s=len(s) / dt, checksum=float(np.sum(s)), d2h_bytes=8 * len(s))
            v2 = g.tc_vertex2()
            emit(config="vertex_count2", scale=args.scale22, sum=int(v2.sum()), consistent=bool(int(v2.sum()) == 6 * tri))
        g.free()
        G.lib().gmsb_trim_memory()

    if 5 in cfgs:
        g, info = build(args.scale26, a=0.65, bc=0.15)
        emit(config="graph", **info)
        best = None
        for rep in range(3):
            tri, st = g.tc_total_ex(reuse_plan=(rep > 0))
            if best is None or st["ms_count"] < best["ms_count"]:
                keep = st["ms_orient"] if rep == 0 else best["ms_orient_first"]
                best = dict(st)
                best["ms_orient_first"] = keep
        emit(config="cfg5-tc-skewed", scale=args.scale26, a=0.65, triangles=tri, count_ms=best["ms_count"],
             orient_ms=best["ms_orient_first"], kernel_ms={"bitmap": best["ms_bitmap"], "merge": best["ms_merge"],
                                                            "gallop": best["ms_gallop"]},
             edges_by_kernel={"bitmap": best["edges_bitmap"], "merge": best["edges_merge"], "gallop": best["edges_gallop"]},
             max_dplus=best["max_dplus"], algorithmic_bytes=best["algorithmic_bytes"],
             edges_per_s=info["m"] / (best["ms_count"] * 1e-3),
             algorithmic_GBps=best["algorithmic_bytes"] / (best["ms_count"] * 1e-3) / 1e9)
        g.free()


if __name__ == "__main__":
    main()
