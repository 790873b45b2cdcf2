"""Phase breakdown of the end-to-end path (host CSR -> HBM -> ranking -> DAG -> schedule -> count) at one scale:

    GMSB_TC_TRACE=1 python tools/e2e_trace.py [--scale 24] [--reps 3]

The library prints the device time of every phase of the representation build on stderr; this script adds the wall
time of the two C-ABI calls (gmsb_graph_from_csr from pinned host memory, gmsb_tc_total_ex with nothing cached)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402  (pinned host memory)
import gms_b200 as G  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=24)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--orient", type=int, default=1, help="1: gmsb_graph_from_csr_ex(GMSB_BUILD_ORIENT)")
    ap.add_argument("--shard-parts", type=int, default=0,
                    help="N > 0: also time the phases of the sharded build (gmsb_shard_*) for part 0 of N on this one "
                         "device; the other parts' pieces are built first, the collectives are not part of the figure")
    args = ap.parse_args()
    G.set_device(0)
    src, dst = G.generate_rmat(args.scale)
    g = G.Graph.from_edgelist(src, dst, True)
    del src, dst
    n, slots = g.n, g.slots
    off_h = torch.empty(n + 1, dtype=torch.int64).pin_memory()
    nbr_h = torch.empty(max(slots, 1), dtype=torch.int32).pin_memory()
    G.capi._check(G.lib().gmsb_graph_export_csr(g.h, off_h.numpy(), nbr_h.numpy()))
    g.free()
    for rep in range(args.reps):
        print(f"--- rep {rep}", file=sys.stderr, flush=True)
        t0 = time.perf_counter()
        gg = G.Graph.from_csr(off_h.numpy(), nbr_h.numpy()[:slots], orient=bool(args.orient))
        t1 = time.perf_counter()
        c, st = gg.tc_total_ex(reuse_plan=False)
        t2 = time.perf_counter()
        gg.free()
        print(json.dumps({"scale": args.scale, "orient_with_upload": args.orient, "rep": rep, "from_csr_ms": round((t1 - t0) * 1e3, 2),
                          "tc_total_ms": round((t2 - t1) * 1e3, 2), "triangles": c,
                          "ms_orient": round(st["ms_orient"], 2), "ms_count": round(st["ms_count"], 2),
                          "ms_bitmap": round(st["ms_bitmap"], 2), "n_items": st["bitmap_items"],
                          "edges": [st["edges_bitmap"], st["edges_merge"], st["edges_gallop"]]}), flush=True)
    if args.shard_parts > 0:
        shard_phases(args, off_h.numpy(), nbr_h.numpy()[:slots], n)


def shard_phases(args, off, nbr, n):
    P = args.shard_parts
    dev = torch.device("cuda", 0)
    others = [G.Shard(off, nbr, p, P) for p in range(1, P)]
    for rep in range(args.reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mine = G.Shard(off, nbr, 0, P)
        t1 = time.perf_counter()
        stride = max([mine.piece_len] + [s.piece_len for s in others])
        pieces = torch.empty(P * stride, dtype=torch.int32, device=dev)
        dplus = torch.zeros(n, dtype=torch.int32, device=dev)
        for p, s in enumerate(others, start=1):
            s.export(pieces.data_ptr() + 4 * p * stride, dplus.data_ptr())
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        mine.export(pieces.data_ptr(), dplus.data_ptr())
        t3 = time.perf_counter()
        gg = mine.finish(pieces.data_ptr(), stride, dplus.data_ptr())
        G.synchronize()
        t4 = time.perf_counter()
        c, st = gg.tc_total_ex(reuse_plan=False, part_index=0, part_count=P)
        t5 = time.perf_counter()
        full = sum(gg.tc_total_ex(reuse_plan=False, part_index=p, part_count=P)[0] for p in range(P)) if rep == 0 else None
        gg.free()
        print(json.dumps({"scale": args.scale, "shard_parts": P, "rep": rep, "piece_len": mine.piece_len, "stride": stride,
                          "shard_begin_ms": round((t1 - t0) * 1e3, 2), "shard_export_ms": round((t3 - t2) * 1e3, 2),
                          "shard_finish_ms": round((t4 - t3) * 1e3, 2), "tc_total_share_ms": round((t5 - t4) * 1e3, 2),
                          "ms_schedule": round(st["ms_orient"], 2), "ms_count": round(st["ms_count"], 2),
                          "triangles_all_parts": full}), flush=True)


if __name__ == "__main__":
    main()
