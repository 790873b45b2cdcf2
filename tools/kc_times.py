"""GPU wall time of gmsb_kclique_count per k on a Kronecker graph, best of <reps> (python tools/kc_times.py <scale> <kmax> [reps])."""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gms_b200 as G
scale, kmax = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
s, d = G.generate_rmat(scale)
g = G.Graph.from_edgelist(s, d, True)
g.kclique_count(3)
for k in range(3, kmax + 1):
    best, c = None, None
    for _ in range(reps):
        t = time.perf_counter()
        c = g.kclique_count(k)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    print(json.dumps({"scale": scale, "k": k, "count": c, "gpu_seconds": best}), flush=True)
