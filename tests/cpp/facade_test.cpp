// C++ facade test: reads like the reference's own tests (testing/set_graph.cpp, testing/clique_counting/*), but the
// graph type is CudaSetGraph.  Exit code 0 = all checks passed; 3 = no usable CUDA device (loud failure).
#include <cstdio>
#include <vector>

#include <gms_b200/gms_api.hpp>

#define EXPECT(cond)                                                             \
    do {                                                                         \
        if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } \
    } while (0)

using namespace GMS;

static CudaSetGraph from_pairs(const std::vector<std::pair<int, int>> &el, bool symmetrize = true) {
    std::vector<NodeId> s, d;
    for (auto &e : el) { s.push_back(e.first); d.push_back(e.second); }
    return CudaSetGraph::FromEdgeList(s.data(), d.data(), (int64_t)s.size(), symmetrize);
}

struct HostCsr {     // a minimal CGraph: what FromCGraph needs (num_nodes, out_degree, out_neigh, directed)
    std::vector<int64_t> off;
    std::vector<NodeId> nbr;
    int64_t num_nodes() const { return (int64_t)off.size() - 1; }
    int64_t out_degree(NodeId v) const { return off[v + 1] - off[v]; }
    gms_b200::NeighborhoodView out_neigh(NodeId v) const { return {nbr.data() + off[v], nbr.data() + off[v + 1]}; }
    bool directed() const { return false; }
};

static int run() {
    // testing/testGraphs/triangles_3.el: TC 3, vertex_count2 [2,4,4,2,0,2,2,2,0,0], degree order
    CudaSetGraph g = from_pairs({{0, 1}, {0, 2}, {1, 2}, {1, 3}, {2, 3}, {5, 6}, {5, 7}, {6, 7}, {6, 8}, {7, 9}, {8, 9}});
    EXPECT(g.num_nodes() == 10 && g.num_edges() == 11 && !g.directed());
    EXPECT(TriangleCount::Par::count_total<CudaSetGraph>(g) == 3);
    EXPECT(TriangleCount::Seq::count_total<CudaSetGraph>(g) == 3);
    std::vector<int64_t> counts;
    TriangleCount::Par::vertex_count2<CudaSetGraph>(g, counts);
    EXPECT((counts == std::vector<int64_t>{2, 4, 4, 2, 0, 2, 2, 2, 0, 0}));
    std::vector<NodeId> order, rank;
    PpParallel::getDegreeOrdering<CudaSetGraph, false>(g, order);
    PpParallel::getDegreeOrdering<CudaSetGraph, true>(g, rank);
    EXPECT((order == std::vector<NodeId>{4, 0, 3, 5, 8, 9, 1, 2, 6, 7}));
    for (int i = 0; i < 10; ++i) EXPECT(rank[order[i]] == i);
    // SGraph concept: out_degree / out_neigh mirror, clone
    EXPECT(g.out_degree(1) == 3 && g.out_neigh(1).cardinality() == 3 && *g.out_neigh(1).begin() == 0);
    CudaSetGraph c = g.clone();
    EXPECT(TriangleCount::Par::count_total<CudaSetGraph>(c) == 3);
    // FromCGraph<CGraph, RemoveIsolated = true> (set_graph.h:190-232, testing/set_graph.cpp): vertex 4 of triangles_3 has
    // no neighbours; it is dropped and 5..9 become 4..8, per-vertex results follow the new ids
    CudaSetGraph shrunk = CudaSetGraph::FromCGraph<CudaSetGraph, true>(g);
    EXPECT(shrunk.num_nodes() == 9 && shrunk.num_edges() == 11);
    EXPECT(TriangleCount::Par::count_total<CudaSetGraph>(shrunk) == 3);
    TriangleCount::Par::vertex_count2<CudaSetGraph>(shrunk, counts);
    EXPECT((counts == std::vector<int64_t>{2, 4, 4, 2, 2, 2, 2, 0, 0}));
    EXPECT(*shrunk.out_neigh(4).begin() == 5);                      // old 5 -> {6, 7}, now 4 -> {5, 6}
    CudaSetGraph same = CudaSetGraph::FromCGraph<CudaSetGraph, true>(shrunk);    // nothing to remove: unchanged
    EXPECT(same.num_nodes() == 9 && TriangleCount::Par::count_total<CudaSetGraph>(same) == 3);
    // testing/set_graph.cpp:69-114 — FromCGraph default / RemoveIsolated on the edge 0-2 (vertex 1 isolated), Clone
    {
        CudaSetGraph with_iso = from_pairs({{0, 2}});
        CudaSetGraph d = CudaSetGraph::FromCGraph(with_iso);
        EXPECT(d.num_nodes() == 3 && d.out_degree(1) == 0 && *d.out_neigh(0).begin() == 2 && *d.out_neigh(2).begin() == 0);
        CudaSetGraph r = CudaSetGraph::FromCGraph<CudaSetGraph, true>(with_iso);
        EXPECT(r.num_nodes() == 2 && *r.out_neigh(0).begin() == 1 && *r.out_neigh(1).begin() == 0);
        CudaSetGraph cl = r.clone();
        EXPECT(cl.num_nodes() == 2 && cl.out_degree(0) == 1 && *cl.out_neigh(1).begin() == 0);
        // FromEL (set_graph.h:54-78): not symmetrised, sorted on request
        std::vector<std::pair<NodeId, NodeId>> el = {{2, 0}, {0, 2}, {0, 1}, {1, 0}, {1, 2}, {2, 1}};
        CudaSetGraph fe = CudaSetGraph::FromEL(el, 3, false);
        EXPECT(fe.num_nodes() == 3 && !fe.directed() && fe.out_degree(0) == 2 && *fe.out_neigh(0).begin() == 1);
        EXPECT(TriangleCount::Par::count_total<CudaSetGraph>(fe) == 1);
        std::vector<std::pair<NodeId, NodeId>> one_way = {{0, 1}, {0, 2}, {1, 2}};
        EXPECT(CudaSetGraph::FromEL(one_way, 3, true).directed());
    }
    // FromCGraph from a host CSR type
    HostCsr h;
    h.off = {0, 2, 4, 6};
    h.nbr = {1, 2, 0, 2, 0, 1};
    CudaSetGraph tri = CudaSetGraph::FromCGraph(h);
    EXPECT(TriangleCount::Par::count_total<CudaSetGraph>(tri) == 1);
    // vertex similarity on triangles_1: Jaccard 1/(2+2+1), CommNeigh 1
    EXPECT((VertexSim::vertex_similarity<VertexSim::Metric::Jaccard>(0, 1, tri) == 0.2));
    EXPECT((VertexSim::vertex_similarity<VertexSim::Metric::CommNeigh>(0, 1, tri) == 1.0));
    auto js = VertexSim::edge_similarity<VertexSim::Metric::Jaccard>(tri);
    EXPECT(js.size() == 3 && js[0] == 0.2 && js[2] == 0.2);
    // k-clique KATs (testing/clique_counting/CliqueCounter2_tests.h:192-224: six 4-cliques)
    CudaSetGraph k4 = from_pairs({{0, 1}, {0, 2}, {0, 3}, {0, 4}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6}, {2, 3}, {2, 4},
                                  {2, 5}, {2, 6}, {3, 4}, {3, 7}, {4, 8}, {5, 6}, {6, 7}, {7, 8}});
    std::vector<NodeId> ranking;
    PpSequential::getDegeneracyOrderingDanischHeap(k4, ranking);
    CudaSetGraph dag = PpSequential::InduceDirectedGraph(k4, ranking);
    EXPECT(dag.directed());
    EXPECT(KClique::Par::EP_kclisting(dag, KClique::CLCliqueApp(4)) == 6);
    EXPECT(KClique::Par::NP_kclisting(dag, KClique::CLCliqueApp(4)) == 6);
    EXPECT(KClique::Seq::Kclisting(dag, KClique::CLCliqueApp(2)) == 19);
    EXPECT(CliqueCount<>(k4, 4) == 6 * 24);
    // InduceDirectedGraph on a directed graph throws std::invalid_argument (apply_order.h:14-16)
    bool thrown = false;
    try { PpSequential::InduceDirectedGraph(dag, ranking); } catch (const std::invalid_argument &) { thrown = true; }
    EXPECT(thrown);
    // batched SortedSet algebra (testing/sets.cpp:133-137): {1..5}∩{3..7} = {3,4,5}
    CudaSetGraph sets = from_pairs({{8, 1}, {8, 2}, {8, 3}, {8, 4}, {8, 5}, {9, 3}, {9, 4}, {9, 5}, {9, 6}, {9, 7}}, false);
    EXPECT(sets.intersect_count({8}, {9})[0] == 3);
    auto is = sets.intersect({8, 9}, {9, 8});
    EXPECT((is.second == std::vector<NodeId>{3, 4, 5, 3, 4, 5}));
    // testing/sets.cpp difference / union KATs on the same sets: {1..5}\{3..7} = {1,2}, union = {1..7}
    auto df = sets.difference({8, 9}, {9, 8});
    EXPECT((df.second == std::vector<NodeId>{1, 2, 6, 7}) && df.first[1] == 2);
    auto un = sets.union_with({8}, {9});
    EXPECT((un.second == std::vector<NodeId>{1, 2, 3, 4, 5, 6, 7}));
    EXPECT(sets.union_count({8, 8}, {9, 8}) == (std::vector<uint64_t>{7, 5}));
    // generated graph: kronecker-12 (golden: 483489 triangles, 4021397 4-cliques)
    CudaSetGraph kron = CudaSetGraph::Kronecker(12);
    EXPECT(TriangleCount::Par::count_total<CudaSetGraph>(kron) == 483489);
    EXPECT(gms_b200::kclique_count(kron, 4) == 4021397ull);
    EXPECT(TriangleCount::Par::count_total<CudaSetGraph>(kron.RelabelByDegree()) == 483489);
    std::printf("facade_test: all checks passed\n");
    return 0;
}

int main() {
    try {
        return run();
    } catch (const gms_b200::Error &e) {
        std::fprintf(stderr, "gms-b200 error %d: %s\n", e.code, e.what());
        return e.code == GMSB_ERR_CUDA ? 3 : 1;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 1;
    }
}
