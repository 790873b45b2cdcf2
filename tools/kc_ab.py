"""A/B timing of clique-kernel settings on one graph: python tools/kc_ab.py <scale> <k> VAR=val[,VAR=val] ... (one run per
argument; 'default' = no override)."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gms_b200 as G
scale, k = int(sys.argv[1]), int(sys.argv[2])
s, d = G.generate_rmat(scale)
g = G.Graph.from_edgelist(s, d, True)
g.kclique_count(3)
for cfg in sys.argv[3:]:
    sets = [] if cfg == "default" else [kv.split("=") for kv in cfg.split(",")]
    for kk, vv in sets:
        os.environ[kk] = vv
    t = time.time()
    c = g.kclique_count(k)
    print(f"k={k} scale={scale} cfg={cfg} count={c} seconds={time.time() - t:.3f}", flush=True)
    for kk, _ in sets:
        del os.environ[kk]
