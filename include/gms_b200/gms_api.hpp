// gms_api.hpp — the reference's algorithm entry points, same names and argument meaning, for CudaSetGraph.
//
// Two ways to use it:
//
//  (1) next to the reference (drop-in).  Include the reference's headers first (triangle_count.h, preprocessing.h)
//      and define GMSB_WITH_GMS_HEADERS (plus GMSB_WITH_GMS_KCLIQUE / GMSB_WITH_GMS_VERTEXSIM when
//      clique_counting.h / vertex_similarity.h are included too): this header then adds EXPLICIT SPECIALISATIONS of the reference's own function
//      templates for SGraph = gms_b200::CudaSetGraph, so existing benchmark mains only change the graph type:
//
//          BenchmarkKernelBk<CudaSetGraph>(args, g, TriangleCount::Par::count_total<CudaSetGraph>,
//                                          TriangleCount::Verify::total_count, "CudaSetGraph");
//
//      (gms/common/benchmark.h:96-137 calls CudaSetGraph::FromCGraph(g) once, then the kernel per trial.)
//
//  (2) stand-alone.  Without the macro this header declares the same namespaces and names itself; the
//      templates accept CudaSetGraph only.
//
// Reference signatures mirrored (file:line in the reference tree):
//   size_t GMS::TriangleCount::{Seq,Par}::count_total(const SGraph&)            triangle_count/parallel/total.h:8
//   void   GMS::TriangleCount::{Seq,Par}::vertex_count2(const SGraph&, Output&) triangle_count/parallel/vertex.h:15
//   void   GMS::TriangleCount::Par::vertex_count2_once(const SGraph&, Output&)  triangle_count/parallel/vertex.h:31
//   void   PpParallel::getDegreeOrdering<G,useRankFormat>(const G&, Output&)    preprocessing/parallel/degree.h:26
//   void   PpSequential::getDegeneracyOrderingDanischHeap(const G&, Output&)    preprocessing/sequential/degeneracy_danisch.h:51
//   G      PpSequential::InduceDirectedGraph(const G&, const std::vector<NodeId>&) preprocessing/sequential/apply_order.h:10
//   ull    GMS::KClique::Par::{NP,EP}_kclisting(G&, clique size)                k_clique_list/clique_counting.h:14-19
//   size_t CliqueCount<Set,SGraph,Set2>(G&, size_t k)                           k_clique_count/k_clique_count_set_based.h:20
//   double GMS::VertexSim::vertex_similarity<Metric>(NodeId, NodeId, const SGraph&) vertex_similarity/vertex_similarity.h:202
#pragma once
#include "cuda_set_graph.hpp"

#ifdef GMSB_WITH_GMS_HEADERS
// ---------------------------------------------------------------------------------------------------------------
// (1) explicit specialisations of the reference's templates
// ---------------------------------------------------------------------------------------------------------------
namespace GMS::TriangleCount::Seq {
template <> inline size_t count_total<gms_b200::CudaSetGraph>(const gms_b200::CudaSetGraph &g) {
    return gms_b200::count_total(g);
}
template <>
inline void vertex_count2<gms_b200::CudaSetGraph, std::vector<int64_t>>(const gms_b200::CudaSetGraph &g,
                                                                        std::vector<int64_t> &counts) {
    gms_b200::vertex_count2(g, counts);
}
}  // namespace GMS::TriangleCount::Seq
namespace GMS::TriangleCount::Par {
template <> inline size_t count_total<gms_b200::CudaSetGraph>(const gms_b200::CudaSetGraph &g) {
    return gms_b200::count_total(g);
}
template <>
inline void vertex_count2<gms_b200::CudaSetGraph, std::vector<int64_t>>(const gms_b200::CudaSetGraph &g,
                                                                        std::vector<int64_t> &counts) {
    gms_b200::vertex_count2(g, counts);
}
template <>
inline void vertex_count2_once<gms_b200::CudaSetGraph, std::vector<int64_t>>(const gms_b200::CudaSetGraph &g,
                                                                             std::vector<int64_t> &counts) {
    gms_b200::vertex_count2(g, counts);
}
}  // namespace GMS::TriangleCount::Par
namespace PpParallel {
template <>
inline void getDegreeOrdering<gms_b200::CudaSetGraph, false, std::vector<NodeId>>(const gms_b200::CudaSetGraph &g,
                                                                                  std::vector<NodeId> &res) {
    gms_b200::degree_ordering<false>(g, res);
}
template <>
inline void getDegreeOrdering<gms_b200::CudaSetGraph, true, std::vector<NodeId>>(const gms_b200::CudaSetGraph &g,
                                                                                 std::vector<NodeId> &res) {
    gms_b200::degree_ordering<true>(g, res);
}
}  // namespace PpParallel
namespace PpSequential {
template <>
inline void getDegeneracyOrderingDanischHeap<gms_b200::CudaSetGraph, std::vector<NodeId>>(
    const gms_b200::CudaSetGraph &g, std::vector<NodeId> &ranking) {
    gms_b200::degeneracy_ordering(g, ranking);
}
template <>
inline gms_b200::CudaSetGraph InduceDirectedGraph<gms_b200::CudaSetGraph>(const gms_b200::CudaSetGraph &g,
                                                                          const std::vector<NodeId> &ranking) {
    return gms_b200::induce_directed_graph(g, ranking);
}
}  // namespace PpSequential
#ifdef GMSB_WITH_GMS_KCLIQUE      // needs gms/algorithms/non_set_based/k_clique_list/clique_counting.h
namespace GMS::KClique::Parallelize {
// Parallelize::{node,edge}<Builder_T, Counter_T, CGraph>(CGraph&, const CLApp&) — clique size read from CLCliqueApp
inline unsigned long long cuda_kclisting(gms_b200::CudaSetGraph &g, const CLApp &cli) {
    const CLCliqueApp &dcli = dynamic_cast<const CLCliqueApp &>(cli);
    return gms_b200::kclique_count(g, dcli.clique_size());
}
}  // namespace GMS::KClique::Parallelize
#endif
#ifdef GMSB_WITH_GMS_VERTEXSIM    // needs gms/algorithms/set_based/vertex_similarity/vertex_similarity.h
namespace GMS::VertexSim {
template <Metric metric>
inline double vertex_similarity(NodeId a, NodeId b, const gms_b200::CudaSetGraph &g) {
    return gms_b200::pair_similarity(static_cast<gms_b200::Metric>(static_cast<int>(metric)), g, {a}, {b})[0];
}
}  // namespace GMS::VertexSim
#endif

#else
// ---------------------------------------------------------------------------------------------------------------
// (2) stand-alone: the same names, declared here
// ---------------------------------------------------------------------------------------------------------------
#include <type_traits>

using NodeId = gms_b200::NodeId;
using CudaSetGraph = gms_b200::CudaSetGraph;

namespace gms_b200::detail {
template <class G> constexpr bool is_cuda_graph = std::is_same_v<std::remove_cv_t<G>, gms_b200::CudaSetGraph>;
}

namespace GMS::TriangleCount::Par {
template <class SGraph> size_t count_total(const SGraph &graph) {
    static_assert(gms_b200::detail::is_cuda_graph<SGraph>, "gms-b200 implements this entry point for CudaSetGraph");
    return gms_b200::count_total(graph);
}
template <class SGraph, class Output = std::vector<int64_t>> void vertex_count2(const SGraph &graph, Output &counts) {
    static_assert(gms_b200::detail::is_cuda_graph<SGraph>, "gms-b200 implements this entry point for CudaSetGraph");
    gms_b200::vertex_count2(graph, counts);
}
template <class SGraph, class Output = std::vector<int64_t>>
void vertex_count2_once(const SGraph &graph, Output &counts) { vertex_count2<SGraph, Output>(graph, counts); }
}  // namespace GMS::TriangleCount::Par
namespace GMS::TriangleCount::Seq {
template <class SGraph> size_t count_total(const SGraph &graph) { return Par::count_total<SGraph>(graph); }
template <class SGraph, class Output = std::vector<int64_t>> void vertex_count2(const SGraph &graph, Output &counts) {
    Par::vertex_count2<SGraph, Output>(graph, counts);
}
}  // namespace GMS::TriangleCount::Seq

namespace PpParallel {
template <class AnyGraph, bool useRankFormat = false, class Output = std::vector<NodeId>>
void getDegreeOrdering(const AnyGraph &graph, Output &res) {
    static_assert(gms_b200::detail::is_cuda_graph<AnyGraph>, "gms-b200 implements this entry point for CudaSetGraph");
    gms_b200::degree_ordering<useRankFormat>(graph, res);
}
// getDegeneracyOrderingApproxCGraph<boundary, useRankFormat>(graph, res, epsilon)   parallel/degeneracy_approx_csr.h:13
// (the averageDegree boundary is the one implemented; the sampled boundaries are randomised variants of it)
namespace boundary_function { struct AverageDegree {}; constexpr AverageDegree averageDegree{}; }
template <const boundary_function::AverageDegree &boundary = boundary_function::averageDegree, bool useRankFormat = false,
          class CGraph = CudaSetGraph, class Output = std::vector<NodeId>>
void getDegeneracyOrderingApproxCGraph(const CGraph &graph, Output &res, double epsilon) {
    static_assert(gms_b200::detail::is_cuda_graph<CGraph>, "gms-b200 implements this entry point for CudaSetGraph");
    gms_b200::degeneracy_ordering_approx<useRankFormat>(graph, res, epsilon);
}
}  // namespace PpParallel
namespace PpSequential {
template <class CGraph = CudaSetGraph, class Output>
void getDegeneracyOrderingDanischHeap(const CGraph &g, Output &ranking) {
    static_assert(gms_b200::detail::is_cuda_graph<CGraph>, "gms-b200 implements this entry point for CudaSetGraph");
    gms_b200::degeneracy_ordering(g, ranking);
}
template <class CGraph = CudaSetGraph>
CGraph InduceDirectedGraph(const CGraph &g, const std::vector<NodeId> &ranking) {
    static_assert(gms_b200::detail::is_cuda_graph<CGraph>, "gms-b200 implements this entry point for CudaSetGraph");
    return gms_b200::induce_directed_graph(g, ranking);      // throws std::invalid_argument on a directed graph
}
}  // namespace PpSequential

namespace GMS::KClique {
// stand-in for CLCliqueApp (parallelizationStrategy/parallelize.h:14-25): carries the clique size
class CLCliqueApp {
public:
    explicit CLCliqueApp(int clique_size = 8) : clique_size_(clique_size) {}
    int clique_size() const { return clique_size_; }
private:
    int clique_size_;
};
namespace Par {
template <class CGraph = CudaSetGraph> unsigned long long EP_kclisting(CGraph &g, const CLCliqueApp &cli) {
    static_assert(gms_b200::detail::is_cuda_graph<CGraph>, "gms-b200 implements this entry point for CudaSetGraph");
    return gms_b200::kclique_count(g, cli.clique_size());
}
template <class CGraph = CudaSetGraph> unsigned long long NP_kclisting(CGraph &g, const CLCliqueApp &cli) {
    return EP_kclisting<CGraph>(g, cli);
}
}  // namespace Par
namespace Seq {
template <class CGraph = CudaSetGraph> unsigned long long Kclisting(CGraph &g, const CLCliqueApp &cli) {
    return Par::EP_kclisting<CGraph>(g, cli);
}
}  // namespace Seq
}  // namespace GMS::KClique

// CliqueCount<Set, SGraph, Set2>(g, k): ordered-tuple convention, returns k! * C_k
template <typename Set = gms_b200::NeighborhoodView, typename SGraph = CudaSetGraph, typename Set2 = Set>
size_t CliqueCount(SGraph &g, size_t k = 4) {
    static_assert(gms_b200::detail::is_cuda_graph<SGraph>, "gms-b200 implements this entry point for CudaSetGraph");
    return gms_b200::kclique_count_ordered(g, k);
}

namespace GMS::VertexSim {
using Metric = gms_b200::Metric;
template <Metric metric, class SGraph> double vertex_similarity(NodeId a, NodeId b, const SGraph &g) {
    static_assert(gms_b200::detail::is_cuda_graph<SGraph>, "gms-b200 implements this entry point for CudaSetGraph");
    return gms_b200::pair_similarity(metric, g, {a}, {b})[0];
}
// batched forms (one kernel launch for all pairs / all edges) — what a caller on the hot path should use
template <Metric metric> std::vector<double> vertex_similarity(const std::vector<NodeId> &a, const std::vector<NodeId> &b,
                                                               const CudaSetGraph &g) {
    return gms_b200::pair_similarity(metric, g, a, b);
}
template <Metric metric> std::vector<double> edge_similarity(const CudaSetGraph &g) {
    return gms_b200::edge_similarity(metric, g);
}
}  // namespace GMS::VertexSim
#endif
