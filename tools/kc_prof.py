import sys, time
sys.path.insert(0, "/root/repo")
import gms_b200 as G
s, d = G.generate_rmat(20)
g = G.Graph.from_edgelist(s, d, True)
t = time.time(); c = g.kclique_count(5); print("k5 s20", c, time.time() - t)
