// TEST INFRASTRUCTURE ONLY — never linked into or called by the product (gms_b200/).
//
// ref_shim.cpp: a thin extern "C" wrapper around the UNMODIFIED reference (spcl/gms) headers.
// It is compiled by oracle/Makefile from the sources where they lie under /root/reference into
// oracle/_ref/libgmsref.so (git-ignored; travels to the GPU box as a prebuilt binary).  It contains no
// reference source text: every function below only *calls* the reference's templates so that
//   (1) oracle/oracle.cpp (our CPU restatement) can be validated against the real thing,
//   (2) tests/golden/ fixtures can be generated from the real thing,
//   (3) bench.py --impl reference / cpu_baseline can time the real thing on the host cores.
//
// Reference entry points that are wrapped (file:line relative to /root/reference):
//   Generator / Builder::MakeGraph            gms/third_party/gapbs/generator.h:81-127, builder.h:1642-1660
//   BuilderBase::MakeGraphFromEL+SquishGraph  gms/third_party/gapbs/builder.h:279-298,237-251
//   BuilderBase::RelabelByDegree              gms/third_party/gapbs/builder.h:1699-1735
//   WorthRelabelling                          gms/third_party/gapbs/benchmark.h:158-176
//   TriangleCount::{Seq,Par}::count_total     gms/algorithms/set_based/triangle_count/{sequential,parallel}/total.h
//   TriangleCount::{Seq,Par}::vertex_count2*  gms/algorithms/set_based/triangle_count/{sequential,parallel}/vertex.h
//   TriangleCount::Verify::compute_total_count gms/algorithms/set_based/triangle_count/verifier.h:14-31
//   PpParallel::getDegreeOrdering             gms/algorithms/preprocessing/parallel/degree.h:26-61
//   PpSequential::getDegeneracyOrderingDanischHeap gms/algorithms/preprocessing/sequential/degeneracy_danisch.h:51-56
//   PpSequential::InduceDirectedGraph         gms/algorithms/preprocessing/sequential/apply_order.h:10-35
//   KClique::{Seq::Kclisting,Par::NP_/EP_kclisting} gms/algorithms/non_set_based/k_clique_list/clique_counting.h:14-34
//   CliqueCount<Set,SGraph,Set2>              gms/algorithms/set_based/k_clique_count/k_clique_count_set_based.h:20-31
//   VertexSim::vertex_similarity<Metric>      gms/algorithms/set_based/vertex_similarity/vertex_similarity.h:202-221
//   SortedSet::{intersect,intersect_count,...} gms/representations/sets/sorted_set.h
#include <unistd.h>
#include <fcntl.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <chrono>
#include <vector>
#include <omp.h>

#include <gms/third_party/gapbs/gapbs.h>
#include <gms/common/cli/cli.h>
#include <gms/representations/graphs/set_graph.h>
#include <gms/representations/sets/sorted_set.h>
#include <gms/algorithms/set_based/triangle_count/triangle_count.h>
#include <gms/algorithms/set_based/triangle_count/verifier.h>
#include <gms/algorithms/set_based/vertex_similarity/vertex_similarity.h>
#include <gms/algorithms/set_based/k_clique_count/k_clique_count_set_based.h>
#include <gms/algorithms/preprocessing/preprocessing.h>
#include <gms/algorithms/non_set_based/k_clique_list/clique_counting.h>

namespace {

// The reference prints progress lines with printf/std::cout from inside library code; silence fd 1 while it runs.
struct Quiet {
    int saved;
    Quiet() {
        fflush(stdout); std::cout.flush();
        saved = dup(1);
        int devnull = open("/dev/null", O_WRONLY);
        dup2(devnull, 1);
        close(devnull);
    }
    ~Quiet() {
        fflush(stdout); std::cout.flush();
        dup2(saved, 1);
        close(saved);
    }
};

struct RefGraph {
    CSRGraph g;
    SortedSetGraph *sg = nullptr;   // lazily built, like BenchmarkKernelBk does once before the trials
    ~RefGraph() { delete sg; }
    const SortedSetGraph &sets() {
        if (!sg) sg = new SortedSetGraph(SortedSetGraph::FromCGraph(g));
        return *sg;
    }
};

GMS::CLI::Args generator_args(int scale, int degree, bool uniform) {
    GMS::CLI::Args args;
    args.graph_spec.is_generator = true;
    args.graph_spec.name = uniform ? "uniform" : "kronecker";
    args.graph_spec.gen_scale = scale;
    args.graph_spec.gen_avgdeg = degree;
    return args;
}

// A CLBase with a chosen symmetrize flag and no generator/file spec.
class ShimCL : public CLApp {
public:
    explicit ShimCL(bool symmetrize) : CLApp(0, nullptr, "shim") { symmetrize_ = symmetrize; }
};

CSRGraph csr_from_arrays(int64_t n, const int64_t *off, const int32_t *nbr, bool directed) {
    int64_t nnz = off[n];
    NodeId *neighs = new NodeId[nnz > 0 ? nnz : 1];
    if (nnz) std::memcpy(neighs, nbr, sizeof(NodeId) * nnz);
    NodeId **index = new NodeId *[n + 1];
    for (int64_t i = 0; i <= n; ++i) index[i] = neighs + off[i];
    if (!directed) return CSRGraph(n, index, neighs);
    // directed graphs in the reference carry an inverse; build it so the destructor's delete[] is well-formed.
    std::vector<int64_t> indeg(n + 1, 0);
    for (int64_t e = 0; e < nnz; ++e) indeg[nbr[e] + 1]++;
    for (int64_t i = 0; i < n; ++i) indeg[i + 1] += indeg[i];
    NodeId *ineighs = new NodeId[nnz > 0 ? nnz : 1];
    NodeId **iindex = new NodeId *[n + 1];
    for (int64_t i = 0; i <= n; ++i) iindex[i] = ineighs + indeg[i];
    std::vector<int64_t> cur(indeg.begin(), indeg.end() - 1);
    for (int64_t u = 0; u < n; ++u)
        for (int64_t e = off[u]; e < off[u + 1]; ++e) ineighs[cur[nbr[e]]++] = (NodeId)u;
    return CSRGraph(n, index, neighs, iindex, ineighs);
}

}  // namespace

extern "C" {

void gmsref_set_threads(int t) { if (t > 0) omp_set_num_threads(t); }
int gmsref_max_threads() { return omp_get_max_threads(); }

// ---- graph construction -------------------------------------------------------------------------------------
void gmsref_generate_el(int scale, int degree, int uniform, int32_t *src, int32_t *dst) {
    Quiet q;
    Generator<NodeId> gen(scale, degree);
    auto el = gen.GenerateEL(uniform != 0);
    int64_t m = (int64_t)el.size();
    #pragma omp parallel for
    for (int64_t e = 0; e < m; ++e) { src[e] = el[e].u; dst[e] = el[e].v; }
}

void *gmsref_generate(int scale, int degree, int uniform) {
    Quiet q;
    auto args = generator_args(scale, degree, uniform != 0);
    auto *h = new RefGraph();
    h->g = args.load_graph();          // Builder(GapbsCompat(args)).MakeGraph(): generate, symmetrise, squish
    return h;
}

void *gmsref_from_el(int64_t m, const int32_t *src, const int32_t *dst, int symmetrize) {
    Quiet q;
    pvector<EdgePair<NodeId, NodeId>> el(m);
    for (int64_t e = 0; e < m; ++e) el[e] = EdgePair<NodeId, NodeId>(src[e], dst[e]);
    ShimCL cl(symmetrize != 0);
    BuilderBase<NodeId, NodeId, NodeId> b(cl);
    auto *h = new RefGraph();
    CSRGraph raw = b.MakeGraphFromEL(el);
    h->g = b.SquishGraph(raw);
    return h;
}

void *gmsref_from_csr(int64_t n, const int64_t *off, const int32_t *nbr, int directed) {
    auto *h = new RefGraph();
    h->g = csr_from_arrays(n, off, nbr, directed != 0);
    return h;
}

void gmsref_free(void *h) { delete static_cast<RefGraph *>(h); }
int64_t gmsref_num_nodes(void *h) { return static_cast<RefGraph *>(h)->g.num_nodes(); }
int64_t gmsref_num_slots(void *h) { return static_cast<RefGraph *>(h)->g.num_edges_directed(); }
int gmsref_directed(void *h) { return static_cast<RefGraph *>(h)->g.directed(); }

void gmsref_export_csr(void *h, int64_t *off, int32_t *nbr) {
    const CSRGraph &g = static_cast<RefGraph *>(h)->g;
    int64_t n = g.num_nodes(), pos = 0;
    for (int64_t u = 0; u < n; ++u) {
        off[u] = pos;
        for (NodeId v : g.out_neigh(u)) nbr[pos++] = v;
    }
    off[n] = pos;
}

// ---- graph files (gapbs reader.h / writer.h through Builder::MakeGraph and WriterBase::WriteGraph) -------------------------
void *gmsref_load_file(const char *path, int symmetrize) {
    Quiet q;
    GMS::CLI::Args args;
    args.graph_spec.is_generator = false;
    args.graph_spec.name = path;
    args.symmetrize = symmetrize != 0;
    auto *h = new RefGraph();
    h->g = args.load_graph();
    return h;
}
void gmsref_write_file(void *h, const char *path, int serialized) {
    Quiet q;
    WriterBase<NodeId> w(static_cast<RefGraph *>(h)->g);
    w.WriteGraph(path, serialized != 0);
}

int gmsref_worth_relabelling(void *h) { return WorthRelabelling(static_cast<RefGraph *>(h)->g) ? 1 : 0; }

void *gmsref_relabel_by_degree(void *h) {
    Quiet q;
    auto *r = new RefGraph();
    r->g = Builder::RelabelByDegree(static_cast<RefGraph *>(h)->g);
    return r;
}

// ---- set algebra ----------------------------------------------------------------------------------------------
uint64_t gmsref_intersect_count(const int32_t *a, int64_t na, const int32_t *b, int64_t nb) {
    SortedSet A(a, (size_t)na), B(b, (size_t)nb);
    return A.intersect_count(B);
}
int64_t gmsref_intersect(const int32_t *a, int64_t na, const int32_t *b, int64_t nb, int32_t *out) {
    SortedSet A(a, (size_t)na), B(b, (size_t)nb);
    SortedSet C = A.intersect(B);
    C.toArray(out);
    return (int64_t)C.cardinality();
}
int64_t gmsref_union(const int32_t *a, int64_t na, const int32_t *b, int64_t nb, int32_t *out) {
    SortedSet A(a, (size_t)na), B(b, (size_t)nb);
    SortedSet C = A.union_with(B);
    if (out) C.toArray(out);
    return (int64_t)C.cardinality();
}
uint64_t gmsref_union_count(const int32_t *a, int64_t na, const int32_t *b, int64_t nb) {
    SortedSet A(a, (size_t)na), B(b, (size_t)nb);
    return A.union_count(B);
}
int64_t gmsref_difference(const int32_t *a, int64_t na, const int32_t *b, int64_t nb, int32_t *out) {
    SortedSet A(a, (size_t)na), B(b, (size_t)nb);
    SortedSet C = A.difference(B);
    C.toArray(out);
    return (int64_t)C.cardinality();
}
int gmsref_contains(const int32_t *a, int64_t na, int32_t x) {
    SortedSet A(a, (size_t)na);
    return A.contains(x) ? 1 : 0;
}

// ---- triangle counting ------------------------------------------------------------------------------------------
uint64_t gmsref_tc_total(void *h, int par) {
    const SortedSetGraph &sg = static_cast<RefGraph *>(h)->sets();
    return par ? GMS::TriangleCount::Par::count_total(sg) : GMS::TriangleCount::Seq::count_total(sg);
}
// returns seconds of the kernel alone (FromCGraph excluded, as BenchmarkKernelBk does)
double gmsref_tc_total_timed(void *h, int par, uint64_t *out) {
    const SortedSetGraph &sg = static_cast<RefGraph *>(h)->sets();
    auto t0 = std::chrono::steady_clock::now();
    *out = par ? GMS::TriangleCount::Par::count_total(sg) : GMS::TriangleCount::Seq::count_total(sg);
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
// variant: 0 Seq::vertex_count2, 1 Par::vertex_count2, 2 Par::vertex_count2_once (output zero-initialised first)
void gmsref_tc_vertex2(void *h, int variant, int64_t *out) {
    RefGraph *r = static_cast<RefGraph *>(h);
    const SortedSetGraph &sg = r->sets();
    std::vector<int64_t> counts;
    if (variant == 0) GMS::TriangleCount::Seq::vertex_count2(sg, counts);
    else if (variant == 1) GMS::TriangleCount::Par::vertex_count2(sg, counts);
    else { counts.assign(r->g.num_nodes(), 0); GMS::TriangleCount::Par::vertex_count2_once(sg, counts); }
    std::memcpy(out, counts.data(), sizeof(int64_t) * counts.size());
}
uint64_t gmsref_tc_verify_total(void *h) {
    return GMS::TriangleCount::Verify::compute_total_count(static_cast<RefGraph *>(h)->g);
}

// Bounded sample of Par::count_total's work for the benchmark's reference arm: every `stride`-th undirected
// edge (u<v, CSR order, starting at `phase`) is intersected with the reference's own
// SortedSet::intersect_count.  Returns seconds; *edges = sampled edges, *sum = Σ|N(u)∩N(v)| over the sample.
double gmsref_tc_total_sample(void *h, int64_t stride, int64_t phase, int64_t *edges, uint64_t *sum) {
    RefGraph *r = static_cast<RefGraph *>(h);
    const SortedSetGraph &sg = r->sets();
    const CSRGraph &g = r->g;
    int64_t n = g.num_nodes();
    std::vector<std::pair<NodeId, NodeId>> picks;
    int64_t idx = 0;
    for (NodeId u = 0; u < n; ++u)
        for (NodeId v : g.out_neigh(u))
            if (u < v) { if (idx % stride == phase) picks.emplace_back(u, v); ++idx; }
    uint64_t total = 0;
    int64_t np = (int64_t)picks.size();
    auto t0 = std::chrono::steady_clock::now();
    #pragma omp parallel for schedule(dynamic, 64) reduction(+:total)
    for (int64_t i = 0; i < np; ++i)
        total += sg.out_neigh(picks[i].first).intersect_count(sg.out_neigh(picks[i].second));
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    *edges = np; *sum = total;
    return dt;
}

// ---- orderings / orientation ----------------------------------------------------------------------------------------
void gmsref_degree_order(void *h, int rank_format, int32_t *out) {
    const CSRGraph &g = static_cast<RefGraph *>(h)->g;
    std::vector<NodeId> res;
    if (rank_format) PpParallel::getDegreeOrdering<CSRGraph, true>(g, res);
    else PpParallel::getDegreeOrdering<CSRGraph, false>(g, res);
    std::memcpy(out, res.data(), sizeof(NodeId) * res.size());
}
void gmsref_degeneracy_danisch_heap(void *h, int32_t *rank_out) {
    const CSRGraph &g = static_cast<RefGraph *>(h)->g;
    std::vector<NodeId> ranking;
    PpSequential::getDegeneracyOrderingDanischHeap(g, ranking);
    std::memcpy(rank_out, ranking.data(), sizeof(NodeId) * ranking.size());
}
void gmsref_adg_order(void *h, double eps, int rank_format, int32_t *out) {
    const CSRGraph &g = static_cast<RefGraph *>(h)->g;
    std::vector<NodeId> res;
    if (rank_format)
        PpParallel::getDegeneracyOrderingApproxCGraph<PpParallel::boundary_function::averageDegree, true>(g, res, eps);
    else
        PpParallel::getDegeneracyOrderingApproxCGraph<PpParallel::boundary_function::averageDegree, false>(g, res, eps);
    std::memcpy(out, res.data(), sizeof(NodeId) * res.size());
}
// boundary: 0 averageDegree, 1 minDegree; pull: 0 = the CSR (push) form, 1 = the Set (pull) form over SortedSetGraph
void gmsref_adg_order_ex(void *h, double eps, int rank_format, int boundary, int pull, int32_t *out) {
    namespace bf = PpParallel::boundary_function;
    const CSRGraph &g = static_cast<RefGraph *>(h)->g;
    std::vector<NodeId> res;
    if (!pull) {
        if (boundary == 0) {
            if (rank_format) PpParallel::getDegeneracyOrderingApproxCGraph<bf::averageDegree, true>(g, res, eps);
            else PpParallel::getDegeneracyOrderingApproxCGraph<bf::averageDegree, false>(g, res, eps);
        } else {
            if (rank_format) PpParallel::getDegeneracyOrderingApproxCGraph<bf::minDegree, true>(g, res, eps);
            else PpParallel::getDegeneracyOrderingApproxCGraph<bf::minDegree, false>(g, res, eps);
        }
    } else {
        SortedSetGraph sg = SortedSetGraph::FromCGraph(g);
        if (boundary == 0) {
            if (rank_format) PpParallel::getDegeneracyOrderingApproxSGraph<bf::averageDegree, true, SortedSetGraph>(sg, res, eps);
            else PpParallel::getDegeneracyOrderingApproxSGraph<bf::averageDegree, false, SortedSetGraph>(sg, res, eps);
        } else {
            if (rank_format) PpParallel::getDegeneracyOrderingApproxSGraph<bf::minDegree, true, SortedSetGraph>(sg, res, eps);
            else PpParallel::getDegeneracyOrderingApproxSGraph<bf::minDegree, false, SortedSetGraph>(sg, res, eps);
        }
    }
    std::memcpy(out, res.data(), sizeof(NodeId) * res.size());
}
void *gmsref_induce_directed(void *h, const int32_t *ranking) {
    Quiet q;
    const CSRGraph &g = static_cast<RefGraph *>(h)->g;
    std::vector<NodeId> rk(ranking, ranking + g.num_nodes());
    auto *r = new RefGraph();
    r->g = PpSequential::InduceDirectedGraph<CSRGraph>(g, rk);
    return r;
}

// ---- k-cliques ----------------------------------------------------------------------------------------------------------
// mode: 0 Seq::Kclisting, 1 Par::NP_kclisting, 2 Par::EP_kclisting; graph must be a DAG from gmsref_induce_directed
uint64_t gmsref_kclique(void *h, int k, int mode) {
    Quiet q;
    CSRGraph &g = static_cast<RefGraph *>(h)->g;
    GMS::CLI::Args args;
    auto value = std::make_shared<std::string>(std::to_string(k));
    GMS::KClique::CLCliqueApp cl(args, GMS::CLI::Param(value));
    if (mode == 0) return GMS::KClique::Seq::Kclisting<CSRGraph>(g, cl);
    if (mode == 1) return GMS::KClique::Par::NP_kclisting<CSRGraph>(g, cl);
    return GMS::KClique::Par::EP_kclisting<CSRGraph>(g, cl);
}
double gmsref_kclique_timed(void *h, int k, int mode, uint64_t *out) {
    auto t0 = std::chrono::steady_clock::now();
    *out = gmsref_kclique(h, k, mode);
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
// set-based CliqueCount on the UNORIENTED graph; returns k! * C_k
uint64_t gmsref_clique_count_set_based(void *h, int k) {
    Quiet q;
    CSRGraph &g = static_cast<RefGraph *>(h)->g;
    return CliqueCount<SortedSet, SortedSetGraph, SortedSet>(g, (size_t)k);
}

// ---- vertex similarity ------------------------------------------------------------------------------------------------------
static double sim_dispatch(int metric, NodeId a, NodeId b, const SortedSetGraph &sg) {
    using namespace GMS::VertexSim;
    switch (metric) {
        case 0: return vertex_similarity<Metric::Jaccard>(a, b, sg);
        case 1: return vertex_similarity<Metric::Overlap>(a, b, sg);
        case 2: return vertex_similarity<Metric::AdamicAdar>(a, b, sg);
        case 3: return vertex_similarity<Metric::Resource>(a, b, sg);
        case 4: return vertex_similarity<Metric::CommNeigh>(a, b, sg);
        case 5: return vertex_similarity<Metric::TotalNeigh>(a, b, sg);
        default: return vertex_similarity<Metric::PrefAtt>(a, b, sg);
    }
}
double gmsref_vertex_similarity(void *h, int metric, int32_t a, int32_t b) {
    return sim_dispatch(metric, a, b, static_cast<RefGraph *>(h)->sets());
}
void gmsref_pair_similarity(void *h, int metric, int64_t npairs, const int32_t *a, const int32_t *b, double *out) {
    const SortedSetGraph &sg = static_cast<RefGraph *>(h)->sets();
    #pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < npairs; ++i) out[i] = sim_dispatch(metric, a[i], b[i], sg);
}
// one score per undirected edge u<v in CSR order (the per-edge driver of BASELINE.json configs[3])
int64_t gmsref_edge_similarity(void *h, int metric, double *out) {
    RefGraph *r = static_cast<RefGraph *>(h);
    const SortedSetGraph &sg = r->sets();
    const CSRGraph &g = r->g;
    int64_t n = g.num_nodes();
    std::vector<int64_t> base(n + 1, 0);
    for (NodeId u = 0; u < n; ++u) {
        int64_t c = 0;
        for (NodeId v : g.out_neigh(u)) c += (u < v);
        base[u + 1] = base[u] + c;
    }
    if (out) {
        #pragma omp parallel for schedule(dynamic, 64)
        for (NodeId u = 0; u < n; ++u) {
            int64_t pos = base[u];
            for (NodeId v : g.out_neigh(u))
                if (u < v) out[pos++] = sim_dispatch(metric, u, v, sg);
        }
    }
    return base[n];
}

}  // extern "C"
