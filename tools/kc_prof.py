"""Profiling helper: one k-clique count on a Kronecker graph (python tools/kc_prof.py <scale> <k>)."""
import sys
import time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gms_b200 as G
scale, k = int(sys.argv[1]), int(sys.argv[2])
s, d = G.generate_rmat(scale)
g = G.Graph.from_edgelist(s, d, True)
t = time.time()
c = g.kclique_count(k)
print(f"k={k} scale={scale} count={c} seconds={time.time() - t:.3f}")
