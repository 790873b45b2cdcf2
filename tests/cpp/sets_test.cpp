// The reference's typed Set tests (testing/sets.cpp:29-489, 54 cases, run there for SortedSet / RoaringSet /
// RobinHoodSet) with the set type swapped for CudaSortedSet, followed by two uses of the Set concept from the
// reference's algorithms: Tomita's pivot rule (maximal_clique_enum/sequential/tomita.h:17-47) and one Bron-Kerbosch
// expansion step, checked against plain host code.  Exit code 0 = all checks passed; 3 = no usable CUDA device.
#include <algorithm>
#include <cstdio>
#include <set>
#include <vector>

#include <gms_b200/cuda_sorted_set.hpp>

using Set = gms_b200::CudaSortedSet;
using gms_b200::CudaSetGraph;
using gms_b200::NodeId;

static int g_fail = 0, g_checks = 0;
#define ASSERT_TRUE(c) do { ++g_checks; if (!(c)) { ++g_fail; std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); } } while (0)
#define ASSERT_FALSE(c) ASSERT_TRUE(!(c))
#define ASSERT_EQ(a, b) ASSERT_TRUE((a) == (b))
#define ASSERT_NE(a, b) ASSERT_TRUE((a) != (b))

static void test_intersect(const Set &a, const Set &b, const Set &expected) {          // sets.cpp:108-115
    ASSERT_EQ(a.intersect(b), expected);
    ASSERT_EQ(b.intersect(a), expected);
    ASSERT_EQ(a.intersect_count(b), expected.cardinality());
    ASSERT_EQ(b.intersect_count(a), expected.cardinality());
}
static void test_intersect_inplace(const Set &a, const Set &b, const Set &expected) {  // sets.cpp:145-158
    const size_t size_a = a.cardinality(), size_b = b.cardinality();
    Set a0 = a.clone();
    a0.intersect_inplace(b);
    ASSERT_EQ(a0, expected);
    Set b1 = b.clone();
    b1.intersect_inplace(a);
    ASSERT_EQ(b1, expected);
    ASSERT_EQ(a.cardinality(), size_a);
    ASSERT_EQ(b.cardinality(), size_b);
}
static void test_union(const Set &a, const Set &b, const Set &expected) {
    ASSERT_EQ(a.union_with(b), expected);
    ASSERT_EQ(b.union_with(a), expected);
}
static void test_union_inplace(const Set &a, const Set &b, const Set &expected) {
    Set a_copy = a.clone(), b_copy = b.clone();
    a_copy.union_inplace(b);
    ASSERT_EQ(a_copy, expected);
    b_copy.union_inplace(a);
    ASSERT_EQ(b_copy, expected);
}
static void test_difference(const Set &a, const Set &b, const Set &left, const Set &right) {
    ASSERT_EQ(a.difference(b), left);
    ASSERT_EQ(b.difference(a), right);
}

static void reference_cases() {
    {   // Equality
        ASSERT_EQ(Set(), Set());
        ASSERT_EQ(Set({2}), Set({2}));
        ASSERT_NE(Set(), Set({2}));
        ASSERT_NE(Set({2}), Set());
        ASSERT_NE(Set({1}), Set({2}));
        ASSERT_EQ(Set({2, 4, 8}), Set({4, 2, 8}));
        ASSERT_NE(Set({4, 8}), Set({4, 2, 8}));
        ASSERT_NE(Set({2, 4, 8}), Set({2, 8}));
        Set a{2, 3, 4}, b{4, 3, 2}, c, d;
        ASSERT_EQ(a, b); ASSERT_EQ(c, d); ASSERT_NE(a, c); ASSERT_NE(d, a);
    }
    {   // Cardinality
        Set a, b{1, 5, 6};
        ASSERT_EQ(a.cardinality(), 0u);
        ASSERT_EQ(b.cardinality(), 3u);
    }
    {   // ConstructVector_Empty / NonEmpty, ConstructPointer_Empty / NonEmpty, ConstructSingleton_Various
        std::vector<NodeId> empty{}, input = {1, 5, 2, 7, 9, 0, 3};
        Set s0(empty);
        ASSERT_EQ(s0.cardinality(), 0u); ASSERT_EQ(s0, Set());
        Set s1(input);
        ASSERT_EQ(s1.cardinality(), 7u); ASSERT_EQ(s1, Set({1, 5, 2, 7, 9, 0, 3}));
        Set s2(empty.data(), 0);
        ASSERT_EQ(s2.cardinality(), 0u); ASSERT_EQ(s2, Set());
        Set s3(input.data(), input.size());
        ASSERT_EQ(s3.cardinality(), 7u); ASSERT_EQ(s3, Set({1, 5, 2, 7, 9, 0, 3}));
        Set set1(5), set2(0);
        ASSERT_EQ(set1, Set({5})); ASSERT_EQ(set1.cardinality(), 1u);
        ASSERT_EQ(set2, Set({0})); ASSERT_EQ(set2.cardinality(), 1u);
    }
    // Intersect_*
    test_intersect(Set(), Set(), Set{});
    test_intersect(Set({1, 2, 3}), Set(), Set{});
    test_intersect(Set({1, 2, 3}), Set({4, 5, 6}), Set({}));
    test_intersect(Set({1, 2, 3, 4, 5}), Set({3, 4, 5, 6, 7}), Set{3, 4, 5});
    test_intersect(Set({1, 2, 3, 4, 5, 6, 7}), Set({2, 4, 6, 8}), Set{2, 4, 6});
    test_intersect(Set({1, 2, 3, 4, 5}), Set({1, 2, 3, 4, 5}), Set{1, 2, 3, 4, 5});
    // IntersectInplace_*
    test_intersect_inplace(Set(), Set(), Set());
    test_intersect(Set({1, 2, 3}), Set(), Set{});
    test_intersect_inplace(Set({1, 2, 3}), Set({4, 5, 6}), Set({}));
    test_intersect_inplace(Set({1, 2, 3, 4, 5}), Set({3, 4, 5, 6, 7}), Set{3, 4, 5});
    test_intersect_inplace(Set({1, 2, 3, 4, 5, 6, 7}), Set({2, 4, 6, 8}), Set{2, 4, 6});
    test_intersect_inplace(Set({1, 2, 3, 4, 5}), Set({1, 2, 3, 4, 5}), Set{1, 2, 3, 4, 5});
    // Union_*
    test_union(Set(), Set(), Set());
    test_union(Set{1, 2, 3}, Set{}, Set{1, 2, 3});
    test_union(Set({1, 2, 3}), Set({4, 5, 6}), Set{1, 2, 3, 4, 5, 6});
    test_union(Set({1, 2, 3, 4, 5}), Set({3, 4, 5, 6, 8}), Set{1, 2, 3, 4, 5, 6, 8});
    test_union(Set({1, 2, 3, 4, 5}), Set({1, 2, 3, 4, 5}), Set{1, 2, 3, 4, 5});
    {   // Union_WithSingleton
        Set set;
        Set res = set.union_with(2).union_with(5).union_with(4).union_with(8).union_with(0).union_with(25);
        ASSERT_EQ(res, Set({2, 5, 4, 8, 0, 25}));
    }
    test_union_inplace(Set(), Set(), Set());
    {   // UnionInplace_Singleton
        Set set;
        set.union_inplace(2);
        ASSERT_EQ(set, Set({2}));
        set.union_inplace(5); set.union_inplace(4); set.union_inplace(8); set.union_inplace(0); set.union_inplace(25);
        ASSERT_EQ(set, Set({2, 5, 4, 8, 0, 25}));
        set.union_inplace(25);
        ASSERT_EQ(set, Set({2, 5, 4, 8, 0, 25}));
    }
    // UnionCount_*
    ASSERT_EQ(Set().union_count(Set()), 0u);
    ASSERT_EQ((Set{1, 2, 3}).union_count(Set{}), 3u);
    ASSERT_EQ(Set({1, 2, 3}).union_count(Set({4, 5, 6})), 6u);
    ASSERT_EQ(Set({1, 2, 3, 4, 5}).union_count(Set({3, 4, 5, 6, 8})), 7u);
    ASSERT_EQ(Set({1, 2, 3, 4, 5}).union_count(Set({1, 2, 3, 4, 5})), 5u);
    // Difference_*
    test_difference(Set(), Set(), Set{}, Set{});
    test_difference(Set({1, 2, 3}), Set(), Set{1, 2, 3}, Set{});
    test_difference(Set({1, 2, 3}), Set({4, 5, 6}), Set{1, 2, 3}, Set{4, 5, 6});
    test_difference(Set({1, 2, 3, 4, 5}), Set({3, 4, 5, 6, 8}), Set{1, 2}, Set{6, 8});
    test_difference(Set({1, 2, 3, 4, 5}), Set({1, 2, 3, 4, 5}), Set{}, Set{});
    {   // Difference_OverlappingVarious
        Set set(std::vector<NodeId>{2, 5, 4, 8, 0, 25});
        ASSERT_EQ(set, Set({2, 5, 4, 8, 0, 25}));
        auto res1 = set.difference(Set(std::vector<NodeId>{1, 2, 3, 4}));
        ASSERT_EQ(res1, Set({5, 8, 0, 25}));
        auto res2 = res1.difference(Set(std::vector<NodeId>{22, 44, 5, 11, 8, 51, 0, 25}));
        ASSERT_EQ(res2, Set());
    }
    {   // Difference_Singleton
        Set set{2, 5, 4, 8, 0, 25};
        Set res = set.difference(2).difference(5).difference(4).difference(25).difference(37).difference(58)
                      .difference(99).difference(0);
        ASSERT_EQ(res, Set({8}));
    }
    {   // DifferenceInplace_Empty / Singleton
        Set e;
        e.difference_inplace(Set());
        ASSERT_EQ(e, Set());
        Set set{2, 5, 4, 8, 0, 25};
        for (NodeId x : {2, 5, 4, 25, 16, 38, 55, 0}) set.difference_inplace(x);
        ASSERT_EQ(set, Set({8}));
    }
    {   // Remove_*
        Set s;
        s.remove(0);
        ASSERT_EQ(s, Set{});
        Set a{1, 2, 3, 4, 5}; a.remove(1); ASSERT_EQ(a, Set({2, 3, 4, 5}));
        Set b{1, 2, 3, 4, 5}; b.remove(3); ASSERT_EQ(b, Set({1, 2, 4, 5}));
        Set c{1, 2, 3, 4, 5}; c.remove(5); ASSERT_EQ(c, Set({1, 2, 3, 4}));
        Set d{1, 2, 4, 5}; d.remove(3); ASSERT_EQ(d, Set({1, 2, 4, 5}));
    }
    {   // Add_*
        Set s; s.add(0); ASSERT_EQ(s, Set({0}));
        Set a{2, 3, 4, 5}; a.add(1); ASSERT_EQ(a, Set({1, 2, 3, 4, 5}));
        Set b{1, 2, 4, 5}; b.add(3); ASSERT_EQ(b, Set({1, 2, 3, 4, 5}));
        Set c{1, 2, 3, 4}; c.add(5); ASSERT_EQ(c, Set({1, 2, 3, 4, 5}));
    }
    {   // Contains_Empty / Various
        Set e;
        for (NodeId x : {7, 12, 88, 1}) ASSERT_FALSE(e.contains(x));
        Set set{2, 5, 4, 8, 0, 25};
        for (NodeId x : {7, 12, 88, 1}) ASSERT_FALSE(set.contains(x));
        for (NodeId x : {5, 4, 2, 8, 0, 25}) ASSERT_TRUE(set.contains(x));
    }
    {   // Range_Empty / Various
        auto r0 = Set::Range(0);
        ASSERT_EQ(r0.cardinality(), 0u); ASSERT_EQ(r0, Set());
        ASSERT_EQ(Set::Range(3), Set({0, 1, 2}));
        ASSERT_EQ(Set::Range(5), Set({0, 1, 2, 3, 4}));
    }
    {   // ToArray_Empty / Basic
        Set empty;
        NodeId data[3] = {42, 42, 42};
        empty.toArray(&data[0]);
        ASSERT_TRUE(data[0] == 42 && data[1] == 42 && data[2] == 42);
        Set set{2, 4, 5};
        std::vector<NodeId> buffer(3, 42);
        set.toArray(buffer.data());
        std::sort(buffer.begin(), buffer.end());
        ASSERT_TRUE((buffer == std::vector<NodeId>{2, 4, 5}));
    }
}

// ---- the Set concept as the reference's algorithms use it --------------------------------------------------------------
// tomitaExample.el (testing/testGraphs): 10 vertices, 15 edges
static const std::vector<std::pair<int, int>> kTomita = {{0, 1}, {1, 2}, {1, 3}, {2, 3}, {2, 4}, {3, 4}, {4, 5}, {4, 6}, {4, 8},
                                                         {5, 6}, {6, 7}, {6, 8}, {6, 9}, {7, 8}, {7, 9}};

static std::vector<std::set<NodeId>> host_adj(int n, const std::vector<std::pair<int, int>> &el) {
    std::vector<std::set<NodeId>> adj(n);
    for (auto &e : el) { adj[e.first].insert(e.second); adj[e.second].insert(e.first); }
    return adj;
}

// Tomita pivot (tomita.h:17-47): the vertex of cand ∪ fini with the most neighbours in cand; ties -> first seen.
static NodeId pivot_through_sets(const CudaSetGraph &g, const Set &cand, const Set &fini) {
    NodeId pivot = -1;
    uint64_t best = 0;
    bool first = true;
    for (const Set *part : {&cand, &fini}) {
        const std::vector<uint64_t> deg = cand.intersect_count_neighbourhoods(g, *part);   // one launch per part
        size_t i = 0;
        for (NodeId v : *part) {
            if (first || deg[i] > best) { pivot = v; best = deg[i]; first = false; }
            ++i;
        }
    }
    return pivot;
}

static void algorithm_cases() {
    std::vector<NodeId> s, d;
    for (auto &e : kTomita) { s.push_back(e.first); d.push_back(e.second); }
    CudaSetGraph g = CudaSetGraph::FromEdgeList(s.data(), d.data(), (int64_t)s.size(), true);
    const auto adj = host_adj(10, kTomita);
    // neighbourhood views equal the host adjacency
    for (NodeId v = 0; v < 10; ++v) {
        Set nv = Set::Neighbourhood(g, v);
        ASSERT_EQ(nv, Set(std::vector<NodeId>(adj[v].begin(), adj[v].end())));
    }
    // pivot over the whole graph, and over a few (cand, fini) splits, against the scalar rule
    for (int split : {10, 7, 4}) {
        std::vector<NodeId> cv, fv;
        for (NodeId v = 0; v < 10; ++v) (v < split ? cv : fv).push_back(v);
        Set cand(cv), fini(fv);
        NodeId want = -1;
        size_t best = 0;
        bool first = true;
        for (const auto *part : {&cv, &fv})
            for (NodeId v : *part) {
                size_t deg = 0;
                for (NodeId c : cv) deg += adj[v].count(c);
                if (first || deg > best) { want = v; best = deg; first = false; }
            }
        ASSERT_EQ(pivot_through_sets(g, cand, fini), want);
        // the scalar form of the same rule, one intersect_count per vertex (tomita.h:17)
        ASSERT_EQ(cand.intersect_count(Set::Neighbourhood(g, want)), best);
    }
    // one Bron-Kerbosch expansion (tomita.h:49-60): for v in cand \ N(pivot):  R+v, cand ∩ N(v), fini ∩ N(v); cand -= v; fini += v
    Set cand = Set::Range(10), fini;
    const NodeId pivot = pivot_through_sets(g, cand, fini);
    Set ext = cand.difference(Set::Neighbourhood(g, pivot));
    std::set<NodeId> hcand, hfini;
    for (NodeId v = 0; v < 10; ++v) hcand.insert(v);
    for (NodeId v : ext) {
        Set nv = Set::Neighbourhood(g, v);
        Set c2 = cand.intersect(nv), f2 = fini.intersect(nv);
        std::vector<NodeId> hc, hf;
        for (NodeId x : hcand) if (adj[v].count(x)) hc.push_back(x);
        for (NodeId x : hfini) if (adj[v].count(x)) hf.push_back(x);
        ASSERT_EQ(c2, Set(hc));
        ASSERT_EQ(f2, Set(hf));
        cand.remove(v); fini.add(v);
        hcand.erase(v); hfini.insert(v);
        ASSERT_EQ(cand, Set(std::vector<NodeId>(hcand.begin(), hcand.end())));
        ASSERT_EQ(fini, Set(std::vector<NodeId>(hfini.begin(), hfini.end())));
    }
    // batched: one set against many device sets, results stay on the device
    Set a{1, 2, 3, 4, 5, 6, 7};
    Set b1{2, 4, 6, 8}, b2, b3{7, 9, 11}, b4{1, 2, 3, 4, 5, 6, 7};
    const auto counts = a.intersect_count_many({&b1, &b2, &b3, &b4});
    ASSERT_TRUE((counts == std::vector<uint64_t>{3, 0, 1, 7}));
    // larger sets: both kernels behind intersect_count (galloping for skewed sizes, merge path for balanced ones)
    std::vector<NodeId> big, third, few = {3, 300, 2999, 3000, 5998, 6001};
    for (NodeId x = 0; x < 6000; x += 2) big.push_back(x);
    for (NodeId x = 0; x < 6000; x += 3) third.push_back(x);
    Set sb(big), st(third), sf(few);
    ASSERT_EQ(sb.intersect_count(st), 1000u);            // multiples of 6 below 6000
    ASSERT_EQ(sb.intersect_count(sf), 3u);               // 300, 3000, 5998
    ASSERT_EQ(sb.union_count(st), 3000u + 2000u - 1000u);
    ASSERT_EQ(sb.difference(st).cardinality(), 2000u);
    ASSERT_EQ(sb.intersect(st).intersect(sf), Set({300, 3000}));
}

int main() {
    try {
        reference_cases();
        algorithm_cases();
    } catch (const gms_b200::Error &e) {
        std::fprintf(stderr, "gms-b200 error %d: %s\n", e.code, e.what());
        return e.code == GMSB_ERR_CUDA ? 3 : 2;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 2;
    }
    std::printf("%d checks, %d failures\n", g_checks, g_fail);
    if (!g_fail) std::printf("all checks passed\n");
    return g_fail ? 1 : 0;
}
