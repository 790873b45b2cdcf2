"""One long k-clique count cut into additive shares (python tools/kc_shares.py <scale> <k> <shares> <budget_s>): every
share is share p of P of the per-vertex sub-problems (gmsb_kclique_count_ex), printed as it finishes; the run stops
early when the first share predicts that the whole count would not fit the time budget."""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gms_b200 as G
scale, k, P, budget = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
s, d = G.generate_rmat(scale)
g = G.Graph.from_edgelist(s, d, True)
m = g.slots // 2
g.kclique_count(3)
total, spent = 0, 0.0
for p in range(P):
    t = time.time()
    c = g.kclique_count(k, p, P)
    sec = time.time() - t
    total += c
    spent += sec
    print(json.dumps({"scale": scale, "k": k, "share": p, "shares": P, "count": c, "seconds": round(sec, 3)}), flush=True)
    if p == 0 and sec * P > budget:
        print(json.dumps({"scale": scale, "k": k, "aborted": f"first share took {sec:.1f}s; {P} shares would exceed {budget}s"}), flush=True)
        sys.exit(2)
print(json.dumps({"scale": scale, "k": k, "count": total, "seconds": round(spent, 3), "cliques_per_s": total / spent,
                  "edges_per_s": m / spent, "n_gpus": 1, "shares": P}), flush=True)
