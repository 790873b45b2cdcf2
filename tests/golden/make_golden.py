"""Generate tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref/libgmsref.so).

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The reference cannot travel to the GPU box, so its outputs are committed here as small fixtures: full arrays for
the tiny graphs of the reference's own test-suite (testing/testGraphs/*.el, restated below as edge lists), and
sha256 digests + totals for generated Kronecker / uniform graphs.

Also restated (inputs and expected answers) and re-verified against the reference while generating:
  * the SortedSet intersect KATs          /root/reference/testing/sets.cpp:108-141
  * the nine k-clique KATs                /root/reference/testing/clique_counting/CliqueCounter2_tests.h:44-269
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as B  # noqa: E402

METRICS = list(B.METRICS)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def el(text):
    pairs = [p.split() for p in text.strip().strip(";").split(";") if p.strip()]
    return [int(a) for a, _ in pairs], [int(b) for _, b in pairs]


# testing/testGraphs/*.el (loaded with symmetrize=true by testing/test_helper.h:12-33)
TEST_GRAPHS = {
    "micro": "0 1;1 0",
    "triangles_1": "0 1;0 2;1 2",
    "triangles_3": "0 1;0 2;1 2;1 3;2 3;5 6;5 7;6 7;6 8;7 9;8 9",
    "smallRandom1": "0 1;0 2;0 6;1 0;1 4;1 5;1 6;1 9;2 0;2 3;2 9;3 2;3 4;4 1;4 3;4 6;4 7;4 8;5 1;5 6;5 8;6 0;6 1;6 4;"
                    "6 5;7 4;7 8;7 9;8 4;8 5;8 7;8 9;9 1;9 2;9 7;9 8",
    "eppsteinExample": "0 1;1 0;1 2;1 3;1 4;1 6;2 1;2 4;2 5;2 8;3 1;3 4;3 7;3 8;4 1;4 2;4 3;4 5;4 6;4 7;4 8;5 2;5 4;"
                       "5 7;6 1;6 4;6 8;7 3;7 4;7 5;7 8;8 2;8 3;8 4;8 6;8 7",
    "tomitaExample": "1 2;1 9;2 1;2 9;2 3;3 2;3 4;3 8;3 9;4 3;4 5;4 6;4 7;4 8;5 4;5 6;6 4;6 5;6 7;6 8;7 4;7 6;7 8;8 3;"
                     "8 4;8 6;8 7;9 1;9 2;9 3",
}

# testing/sets.cpp:108-141 — (a, b, a∩b)
SET_KATS = [
    ([], [], []),
    ([1, 2, 3], [], []),
    ([1, 2, 3], [4, 5, 6], []),
    ([1, 2, 3, 4, 5], [3, 4, 5, 6, 7], [3, 4, 5]),
    ([1, 2, 3, 4, 5, 6, 7], [2, 4, 6, 8], [2, 4, 6]),
    ([1, 2, 3, 4, 5], [1, 2, 3, 4, 5], [1, 2, 3, 4, 5]),
]

# testing/clique_counting/CliqueCounter2_tests.h:44-269 — (edges, k, expected)
_HEX = "1 2;2 3;3 4;4 5;5 6;6 1;0 1;0 2;0 3;0 4;0 5;0 6"
_K4A = "0 1;0 2;0 3;0 4;1 2;1 3;1 4;1 5;1 6;2 3;2 4;2 5;2 6;3 4;3 7;4 8;5 6;6 7;7 8"
_K4B = ("0 1;0 2;0 3;1 2;1 3;2 3;2 8;3 12;4 5;4 6;4 7;4 9;5 6;5 7;6 7;6 12;7 13;8 9;8 10;8 11;9 10;9 11;10 11;11 14;"
        "12 13;12 14;12 15;13 14;13 15;14 15")
CLIQUE_KATS = [
    ("0 1;1 2", 2, 2), ("0 1;1 2;0 3;0 4;0 5", 2, 5), ("0 1;1 2", 3, 0), ("0 1;1 2;0 3;0 4;0 5", 3, 0),
    ("0 1;1 2;2 0", 3, 1), (_HEX, 3, 6), (_HEX, 4, 0), (_K4A, 4, 6), (_K4B, 4, 4),
]


def graph_record(R, g, full, kmax, setk):
    off, nbr = g.csr()
    rec = {"n": g.n, "slots": g.slots, "csr_sha": sha(off) + sha(nbr), "tc": g.tc_total(True),
           "worth_relabelling": g.worth_relabelling()}
    assert g.tc_total(False) == rec["tc"] == g.tc_verify_total()
    v2 = g.tc_vertex2(1)
    assert (v2 == g.tc_vertex2(0)).all() and (v2 == g.tc_vertex2(2)).all()
    order, rank = g.degree_order(False), g.degree_order(True)
    dag = g.induce_directed(rank)
    doff, dnbr = dag.csr()
    rec.update({"vertex2_sha": sha(v2), "order_sha": sha(order), "rank_sha": sha(rank), "dag_n": dag.n,
                "dag_sha": sha(doff) + sha(dnbr)})
    rec["kclique"] = {}
    for k in range(1, kmax + 1):
        c = dag.kclique(k, 2)
        assert c == dag.kclique(k, 1) == dag.kclique(k, 0)
        rec["kclique"][str(k)] = c
    rec["ordered"] = {str(k): g.clique_count_set_based(k) for k in setk}
    rel = g.relabel_by_degree()
    ro, rn = rel.csr()
    rec["relabel_sha"] = sha(ro) + sha(rn)
    rec["sim_sha"] = {m: sha(g.edge_similarity(m)) for m in METRICS}
    if full:
        rec.update({"off": off.tolist(), "nbr": nbr.tolist(), "vertex2": v2.tolist(), "order": order.tolist(),
                    "rank": rank.tolist(), "dag_off": doff.tolist(), "dag_nbr": dnbr.tolist(),
                    "relabel_off": ro.tolist(), "relabel_nbr": rn.tolist(),
                    "sim_hex": {m: [x.hex() for x in g.edge_similarity(m)] for m in METRICS}})
    return rec


def main():
    R = B.reference()
    assert R is not None, "build oracle/_ref/libgmsref.so first (make -C oracle)"
    out = {"sets": [], "clique_kats": [], "graphs": {}, "generated": {}}

    for a, b, want in SET_KATS:
        assert R.intersect(a, b).tolist() == want and R.intersect(b, a).tolist() == want
        assert R.intersect_count(a, b) == len(want) == R.intersect_count(b, a)
        out["sets"].append({"a": a, "b": b, "intersect": want, "union": R.union(a, b).tolist(),
                            "difference": R.difference(a, b).tolist(), "union_count": R.union_count(a, b)})

    for text, k, want in CLIQUE_KATS:
        s, d = el(text)
        g = R.from_el(s, d, True)
        rank = g.degree_order(True)
        dag = g.induce_directed(rank)
        assert dag.kclique(k, 0) == dag.kclique(k, 1) == dag.kclique(k, 2) == want, (text, k)
        # the reference's own fixture orients by its heap degeneracy order; the count must agree
        dag2 = g.induce_directed(R.degeneracy_rank(g))
        assert dag2.kclique(k, 0) == want
        out["clique_kats"].append({"src": s, "dst": d, "k": k, "count": want})

    for name, text in TEST_GRAPHS.items():
        s, d = el(text)
        g = R.from_el(s, d, True)
        rec = graph_record(R, g, True, 6, (3, 4, 5))
        rec.update({"src": s, "dst": d})
        out["graphs"][name] = rec

    for kind, scale, kmax, setk in (("kronecker", 8, 6, (3, 4)), ("kronecker", 10, 6, (3, 4)), ("kronecker", 12, 5, (3,)),
                                    ("uniform", 10, 5, (3, 4)), ("kronecker", 14, 4, ())):
        g = R.generate(scale, 16, kind == "uniform")
        rec = graph_record(R, g, False, kmax, setk)
        s, d = R.generate_el(scale, 16, kind == "uniform")
        rec["el_sha"] = sha(s) + sha(d)
        out["generated"][f"{kind}-{scale}"] = rec
    # totals quoted in SURVEY.md §8c, re-measured here
    g = R.generate(16, 16, False)
    out["generated"]["kronecker-16"] = {"n": g.n, "slots": g.slots, "tc": g.tc_total(True)}
    assert out["generated"]["kronecker-16"]["tc"] == 15656287 and out["generated"]["kronecker-12"]["tc"] == 483489

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
