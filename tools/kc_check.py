"""GPU check of the lane-parallel clique kernels against closed forms, the goldens, the oracle and the warp kernels,
followed by timings of both kernel families (python tools/kc_check.py [--scale 22] [--kmax-lane 6] [--kmax-warp 5]).
One JSON line per check / timing; exit code 1 on any mismatch."""
import argparse
import json
import os
import sys
import time
from math import comb

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gms_b200 as G  # noqa: E402

BAD = 0


def emit(**kw):
    print(json.dumps(kw), flush=True)


def count(g, k, impl, parts=1):
    os.environ["GMSB_KCLIQUE_IMPL"] = impl
    t0 = time.time()
    c = sum(g.kclique_count(k, p, parts) for p in range(parts))
    return c, time.time() - t0


def check(name, g, k, want, parts=1):
    global BAD
    got, sec = count(g, k, "lane", parts)
    ok = got == want
    BAD += 0 if ok else 1
    emit(check=name, k=k, parts=parts, ok=ok, got=got, want=want, seconds=round(sec, 3))


def complete(n):
    i, j = np.triu_indices(n, 1)
    return G.Graph.from_edgelist(i.astype(np.int32), j.astype(np.int32), True)


def main():
    global BAD
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--kmax-lane", type=int, default=6)
    ap.add_argument("--kmax-warp", type=int, default=5)
    ap.add_argument("--skip-checks", action="store_true")
    args = ap.parse_args()
    golden = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "golden.json")))

    if not args.skip_checks:
        # closed forms: every class of the lane kernels (d+ <= 64, 128, 256, 512, > 512 in shared memory, > 512 with the
        # CTA-wide levels, matrix spilled to global memory)
        for n, ks in ((40, (4, 5, 6, 7, 8)), (100, (4, 5, 6, 7)), (200, (4, 5, 6)), (300, (4, 5)), (600, (4, 5, 6)),
                      (700, (4, 5)), (1100, (4, 5)), (1300, (4,))):
            g = complete(n)
            for k in ks:
                check(f"K_{n}", g, k, comb(n, k), parts=3 if n in (100, 600) else 1)
            g.free()
        for key in ("kronecker-12", "kronecker-14", "kronecker-16"):
            rec = golden["generated"].get(key, {}).get("kclique")
            if not rec:
                continue
            s, d = G.generate_rmat(int(key.split("-")[1]))
            g = G.Graph.from_edgelist(s, d, True)
            for k in sorted(int(x) for x in rec):
                if 4 <= k <= 10:
                    check(key, g, k, rec[str(k)], parts=2 if k == 5 else 1)
            g.free()
        # dense random graphs against the warp kernels (themselves checked against the oracle by the test-suite)
        rng = np.random.default_rng(7)
        for n, p in ((500, 0.5), (900, 0.35), (1500, 0.6)):
            a = np.triu(rng.random((n, n)) < p, 1)
            i, j = np.nonzero(a)
            g = G.Graph.from_edgelist(i.astype(np.int32), j.astype(np.int32), True)
            for k in (4, 5, 6):
                want, _ = count(g, k, "warp")
                check(f"G({n},{p})", g, k, want)
            g.free()
        emit(summary="checks", mismatches=BAD)

    s, d = G.generate_rmat(args.scale)
    g = G.Graph.from_edgelist(s, d, True)
    m = g.slots // 2
    g.kclique_count(3)
    results = {}
    for k in range(4, max(args.kmax_lane, args.kmax_warp) + 1):
        for impl, kmax in (("lane", args.kmax_lane), ("warp", args.kmax_warp)):
            if k > kmax:
                continue
            c, sec = count(g, k, impl)
            results[(impl, k)] = c
            emit(timing=impl, scale=args.scale, k=k, count=c, seconds=round(sec, 4), cliques_per_s=c / sec, edges_per_s=m / sec)
        if ("lane", k) in results and ("warp", k) in results and results[("lane", k)] != results[("warp", k)]:
            BAD += 1
            emit(check="lane-vs-warp", k=k, ok=False)
    return 1 if BAD else 0


if __name__ == "__main__":
    sys.exit(main())
