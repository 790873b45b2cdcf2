"""Algorithmic bytes B_k of the set-algebra clique recursion, counted exactly by the instrumented CPU
restatement (oracle/oracle.cpp: dag_clique_bytes) on the degree-oriented DAG of a Kronecker graph:

    python tools/kclique_bytes.py <scale> <kmax>        # CPU only; one JSON line per k

Every intersection S ∩ N+(v) of the recursion counts 4·(|S| + d+(v)) bytes; B_3 equals B_TC.  The GPU kernels do not
stream these bytes — they build one bit matrix per vertex and search it on chip — so this is the yardstick the
reference formulation would be held to, reported next to the measured GPU times in DESIGN.md §3."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import binding  # noqa: E402  (test infrastructure: this tool measures the oracle, not the product)

scale, kmax = int(sys.argv[1]), int(sys.argv[2])
binding.build()
o = binding.oracle()
g = o.generate(scale)
dag = g.induce_directed(g.degree_order(True))
for k in range(3, kmax + 1):
    t = time.time()
    b, c = o.kclique_bytes(dag, k)
    print(json.dumps({"scale": scale, "k": k, "algorithmic_bytes": b, "count": c, "cpu_seconds": round(time.time() - t, 2),
                      "cpu_threads": o.max_threads()}), flush=True)
