"""GPU parity (-m gpu): the multi-GPU partition of the clique kernels is additive."""
import pytest

from conftest import random_graph_edges

pytestmark = pytest.mark.gpu


def test_clique_partition_sums_to_total(gms, golden):
    s, d = gms.generate_rmat(12)
    g = gms.Graph.from_edgelist(s, d, True)
    rec = golden["generated"]["kronecker-12"]["kclique"]
    for k in (1, 2, 3, 4, 5):
        for parts in (2, 3, 8):
            shares = [g.kclique_count(k, p, parts) for p in range(parts)]
            assert sum(shares) == rec[str(k)], (k, parts)
    dag = g.orient(g.degree_order(True))
    assert sum(dag.kclique_count(4, p, 4) for p in range(4)) == rec["4"]
    with pytest.raises(gms.GmsbError):
        g.kclique_count(4, 3, 3)


def test_clique_partition_on_dense_graph(gms, orc):
    s, d = random_graph_edges(61, 600, 60000, skew=0.5)
    g, o = gms.Graph.from_edgelist(s, d, True), orc.from_el(s, d, True)
    want = o.induce_directed(o.degree_order(True)).kclique(4)
    assert sum(g.kclique_count(4, p, 5) for p in range(5)) == want
