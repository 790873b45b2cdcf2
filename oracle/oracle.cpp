// TEST INFRASTRUCTURE ONLY — never linked into or called by the product (gms_b200/).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
//
// oracle.cpp: an independent CPU restatement of the reference's (spcl/gms) algorithms for the
// set-intersection hot path, written from the behaviour described at the cited reference lines
// (paths relative to /root/reference).  PARITY IS PINNED: tests/test_oracle_golden.py checks every
// function below against the unmodified reference (oracle/_ref/libgmsref.so, built from the reference's own
// sources by oracle/Makefile) when that library is present, and tests/golden/*.json holds vectors generated
// from the reference itself (tests/golden/make_golden.py) plus the reference's own KATs
// (testing/sets.cpp:108-141, testing/clique_counting/CliqueCounter2_tests.h:44-269).
//
// Everything is plain integer work on int32 ids / int64 offsets, except the similarity scores (IEEE double).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <random>
#include <utility>
#include <vector>
#include <omp.h>

namespace {

using vid = int32_t;
constexpr int64_t kSeed = 27491095;          // gms/third_party/gapbs/util.h:25
constexpr int64_t kGenBlock = int64_t(1) << 18;  // gms/third_party/gapbs/generator.h:140

struct Graph {
    int64_t n = 0;
    bool directed = false;
    std::vector<int64_t> off;   // n+1
    std::vector<vid> nbr;       // off[n]
    int64_t deg(vid v) const { return off[v + 1] - off[v]; }
    const vid *begin(vid v) const { return nbr.data() + off[v]; }
    const vid *end(vid v) const { return nbr.data() + off[v + 1]; }
};

// ---- set algebra on ascending, duplicate-free id ranges ---------------------------------------------------
// |A∩B| by a two-pointer walk (gms/representations/sets/sorted_set_operations.h:45-71).
inline uint64_t isect_count(const vid *a, const vid *ae, const vid *b, const vid *be) {
    uint64_t c = 0;
    while (a != ae && b != be) {
        vid x = *a, y = *b;
        if (x == y) { ++c; ++a; ++b; }
        else if (x > y) ++b;
        else ++a;
    }
    return c;
}
// A∩B materialised (std::set_intersection semantics, sorted_set_operations.h:37-42).
inline int64_t isect_write(const vid *a, const vid *ae, const vid *b, const vid *be, vid *out) {
    vid *o = out;
    while (a != ae && b != be) {
        if (*a < *b) ++a;
        else if (*b < *a) ++b;
        else { *o++ = *a; ++a; ++b; }
    }
    return o - out;
}
// |A∪B| (gms/representations/sets/sorted_set.h:139-158).
inline uint64_t union_count(const vid *a, const vid *ae, const vid *b, const vid *be) {
    uint64_t c = 0;
    while (a != ae && b != be) {
        ++c;
        if (*a == *b) { ++a; ++b; }
        else if (*a > *b) ++b;
        else ++a;
    }
    return c + (ae - a) + (be - b);
}

// ---- generator ------------------------------------------------------------------------------------------------
// R-MAT with per-block reseeding and a final id permutation (gms/third_party/gapbs/generator.h:81-114,52-62).
// The reference hard-codes A=.57 B=.19 C=.19 ("kronecker"); BASELINE.json configs[4] needs A=.65, B=C=.15.
// Bit-exactness with the reference relies on using the same libstdc++ <random> (mt19937, float
// uniform_real_distribution, std::shuffle) — SURVEY.md §7 "Bit-exact inputs".
void rmat_edges(int scale, int64_t m, float A, float B, float C, bool permute, vid *src, vid *dst) {
    const float ab = A + B, abc = A + B + C;
    #pragma omp parallel
    {
        std::mt19937 rng;
        std::uniform_real_distribution<float> unit(0, 1.0f);
        #pragma omp for
        for (int64_t blk = 0; blk < m; blk += kGenBlock) {
            rng.seed(kSeed + blk / kGenBlock);
            int64_t hi = std::min(blk + kGenBlock, m);
            for (int64_t e = blk; e < hi; ++e) {
                vid s = 0, d = 0;
                for (int lvl = 0; lvl < scale; ++lvl) {
                    float p = unit(rng);
                    s <<= 1; d <<= 1;
                    if (p < ab) { if (p > A) d++; }
                    else { s++; if (p > abc) d++; }
                }
                src[e] = s; dst[e] = d;
            }
        }
    }
    if (permute) {
        int64_t n = int64_t(1) << scale;
        std::vector<vid> perm(n);
        std::iota(perm.begin(), perm.end(), 0);
        std::mt19937 rng(kSeed);
        std::shuffle(perm.begin(), perm.end(), rng);
        #pragma omp parallel for
        for (int64_t e = 0; e < m; ++e) { src[e] = perm[src[e]]; dst[e] = perm[dst[e]]; }
    }
}
// uniform random endpoints (generator.h:64-79)
void uniform_edges(int scale, int64_t m, vid *src, vid *dst) {
    int64_t n = int64_t(1) << scale;
    #pragma omp parallel
    {
        std::mt19937 rng;
        std::uniform_int_distribution<vid> pick(0, (vid)(n - 1));
        #pragma omp for
        for (int64_t blk = 0; blk < m; blk += kGenBlock) {
            rng.seed(kSeed + blk / kGenBlock);
            int64_t hi = std::min(blk + kGenBlock, m);
            for (int64_t e = blk; e < hi; ++e) {
                // Edge(udist(rng), udist(rng)): g++ evaluates constructor arguments right to left, so the FIRST
                // draw is the destination (generator.h:74)
                dst[e] = pick(rng); src[e] = pick(rng);
            }
        }
    }
}

// ---- edge list -> squished CSR -------------------------------------------------------------------------------------
// MakeGraphFromEL + SquishGraph (gms/third_party/gapbs/builder.h:279-298,206-251): n = max id + 1, both
// directions inserted when symmetrising, then every list sorted, de-duplicated and stripped of the self loop.
Graph build_from_edges(int64_t m, const vid *src, const vid *dst, bool symmetrize) {
    Graph g;
    vid mx = 0;
    for (int64_t e = 0; e < m; ++e) mx = std::max(mx, std::max(src[e], dst[e]));
    g.n = int64_t(mx) + 1;
    g.directed = !symmetrize;
    std::vector<int64_t> cnt(g.n + 1, 0);
    for (int64_t e = 0; e < m; ++e) { cnt[src[e] + 1]++; if (symmetrize) cnt[dst[e] + 1]++; }
    for (int64_t i = 0; i < g.n; ++i) cnt[i + 1] += cnt[i];
    std::vector<vid> raw(cnt[g.n]);
    {
        std::vector<int64_t> cur(cnt.begin(), cnt.end() - 1);
        for (int64_t e = 0; e < m; ++e) {
            raw[cur[src[e]]++] = dst[e];
            if (symmetrize) raw[cur[dst[e]]++] = src[e];
        }
    }
    std::vector<int64_t> keep(g.n);
    #pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t u = 0; u < g.n; ++u) {
        vid *b = raw.data() + cnt[u], *e = raw.data() + cnt[u + 1];
        std::sort(b, e);
        e = std::unique(b, e);
        e = std::remove(b, e, (vid)u);
        keep[u] = e - b;
    }
    g.off.assign(g.n + 1, 0);
    for (int64_t u = 0; u < g.n; ++u) g.off[u + 1] = g.off[u] + keep[u];
    g.nbr.resize(g.off[g.n]);
    #pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t u = 0; u < g.n; ++u)
        std::copy(raw.data() + cnt[u], raw.data() + cnt[u] + keep[u], g.nbr.data() + g.off[u]);
    return g;
}

// RelabelByDegree (builder.h:1699-1735): new id = position in (degree desc, id desc) order.
Graph relabel_by_degree(const Graph &g) {
    std::vector<std::pair<int64_t, vid>> key(g.n);
    for (int64_t v = 0; v < g.n; ++v) key[v] = {g.deg((vid)v), (vid)v};
    std::sort(key.begin(), key.end(), std::greater<std::pair<int64_t, vid>>());
    std::vector<vid> newid(g.n);
    Graph r; r.n = g.n; r.directed = false; r.off.assign(g.n + 1, 0);
    for (int64_t i = 0; i < g.n; ++i) { newid[key[i].second] = (vid)i; r.off[i + 1] = r.off[i] + key[i].first; }
    r.nbr.resize(r.off[g.n]);
    #pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t u = 0; u < g.n; ++u) {
        vid *o = r.nbr.data() + r.off[newid[u]];
        int64_t k = 0;
        for (const vid *p = g.begin((vid)u); p != g.end((vid)u); ++p) o[k++] = newid[*p];
        std::sort(o, o + k);
    }
    return r;
}

// WorthRelabelling (gms/third_party/gapbs/benchmark.h:158-176, SourcePicker :51-75).
bool worth_relabelling(const Graph &g) {
    int64_t num_edges = g.directed ? g.off[g.n] : g.off[g.n] / 2;      // CSRGraph::num_edges()
    if (num_edges / g.n < 10) return false;
    std::mt19937 rng(kSeed);
    std::uniform_int_distribution<vid> pick(0, (vid)(g.n - 1));
    int64_t ns = std::min<int64_t>(1000, g.n), total = 0;
    std::vector<int64_t> s(ns);
    for (int64_t t = 0; t < ns; ++t) {
        vid v;
        do { v = pick(rng); } while (g.deg(v) == 0);
        s[t] = g.deg(v); total += s[t];
    }
    std::sort(s.begin(), s.end());
    double avg = double(total) / ns, med = double(s[ns / 2]);
    return avg / 1.3 > med;
}

// ---- orderings / orientation ------------------------------------------------------------------------------------------
// PpParallel::getDegreeOrdering (gms/algorithms/preprocessing/parallel/degree.h:16-61): total order
// (degree asc, id asc); order format res[i] = vertex at position i, rank format res[v] = position of v.
void degree_order(const Graph &g, bool rank_format, vid *out) {
    std::vector<vid> ord(g.n);
    std::iota(ord.begin(), ord.end(), 0);
    std::sort(ord.begin(), ord.end(), [&](vid a, vid b) {
        int64_t da = g.deg(a), db = g.deg(b);
        return da < db || (da == db && a < b);
    });
    if (rank_format) for (int64_t i = 0; i < g.n; ++i) out[ord[i]] = (vid)i;
    else std::copy(ord.begin(), ord.end(), out);
}

// Degeneracy ranking with the reference's convention (degeneracy_danisch.h:12-49): repeatedly remove a vertex
// of minimum residual degree; the r-th removed vertex (r = 1..n) gets ranking n - r.  The reference's
// tie-breaking depends on std::make_heap and scatter order and is not reproducible (its own test says so,
// testing/preprocessing.cpp:6-7) — so this is pinned by VALIDITY (orc_check_degeneracy_rank), not by
// permutation.  Ties here: lowest residual degree, most recently bucketed vertex first.
void degeneracy_rank(const Graph &g, vid *rank) {
    int64_t n = g.n;
    std::vector<int64_t> d(n);
    int64_t maxd = 0;
    for (int64_t v = 0; v < n; ++v) { d[v] = g.deg((vid)v); maxd = std::max(maxd, d[v]); }
    std::vector<std::vector<vid>> bucket(maxd + 1);
    for (int64_t v = n - 1; v >= 0; --v) bucket[d[v]].push_back((vid)v);
    std::vector<char> gone(n, 0);
    int64_t removed = 0, cur = 0;
    while (removed < n) {
        while (cur <= maxd && bucket[cur].empty()) ++cur;
        vid v = bucket[cur].back(); bucket[cur].pop_back();
        if (gone[v] || d[v] != cur) continue;      // stale entry
        gone[v] = 1;
        rank[v] = (vid)(n - (++removed));
        for (const vid *p = g.begin(v); p != g.end(v); ++p) {
            vid w = *p;
            if (!gone[w]) { --d[w]; bucket[d[w]].push_back(w); if (d[w] < cur) cur = d[w]; }
        }
    }
}
// Valid iff, processing vertices by DESCENDING rank (= removal order), each vertex has minimum residual degree
// among those not yet removed.  Returns the degeneracy (max residual degree at removal) or -1 if invalid.
int64_t check_degeneracy_rank(const Graph &g, const vid *rank) {
    int64_t n = g.n;
    std::vector<vid> at(n, -1);
    for (int64_t v = 0; v < n; ++v) { if (rank[v] < 0 || rank[v] >= n || at[rank[v]] != -1) return -1; at[rank[v]] = (vid)v; }
    std::vector<int64_t> d(n), hist;
    int64_t maxd = 0;
    for (int64_t v = 0; v < n; ++v) { d[v] = g.deg((vid)v); maxd = std::max(maxd, d[v]); }
    hist.assign(maxd + 2, 0);
    for (int64_t v = 0; v < n; ++v) hist[d[v]]++;
    int64_t degen = 0, lo = 0;
    for (int64_t r = n - 1; r >= 0; --r) {
        vid v = at[r];
        while (hist[lo] == 0) ++lo;
        if (d[v] != lo) return -1;
        degen = std::max(degen, d[v]);
        hist[d[v]]--; d[v] = -1;
        for (const vid *p = g.begin(v); p != g.end(v); ++p)
            if (d[*p] >= 0) { hist[d[*p]]--; d[*p]--; hist[d[*p]]++; if (d[*p] < lo) lo = d[*p]; }
    }
    return degen;
}

// InduceDirectedGraph (gms/algorithms/preprocessing/sequential/apply_order.h:10-35): relabel by ranking, keep
// rank(u) < rank(v); the rebuilt CSR has n = max surviving id + 1 (builder.h:285) and id-sorted lists.
Graph induce_directed(const Graph &g, const vid *rank) {
    std::vector<vid> s, d;
    for (int64_t u = 0; u < g.n; ++u)
        for (const vid *p = g.begin((vid)u); p != g.end((vid)u); ++p)
            if (rank[u] < rank[*p]) { s.push_back(rank[u]); d.push_back(rank[*p]); }
    Graph r = build_from_edges((int64_t)s.size(), s.data(), d.data(), false);
    return r;
}

// ---- triangle counting --------------------------------------------------------------------------------------------------
// count_total (triangle_count/{sequential,parallel}/total.h:8-24): unoriented, u<v pairs, sum / 3.
uint64_t tc_total(const Graph &g, bool par) {
    uint64_t total = 0;
    #pragma omp parallel for schedule(static, 17) reduction(+:total) if (par)
    for (int64_t u = 0; u < g.n; ++u)
        for (const vid *p = g.begin((vid)u); p != g.end((vid)u); ++p)
            if (u < *p) total += isect_count(g.begin((vid)u), g.end((vid)u), g.begin(*p), g.end(*p));
    return total / 3;
}
// vertex_count2 (triangle_count/parallel/vertex.h:15-27): counts[u] = Σ_{v∈N(u)} |N(u)∩N(v)| = 2·t(u).
void tc_vertex2(const Graph &g, int64_t *out) {
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t u = 0; u < g.n; ++u) {
        int64_t c = 0;
        for (const vid *p = g.begin((vid)u); p != g.end((vid)u); ++p)
            c += (int64_t)isect_count(g.begin((vid)u), g.end((vid)u), g.begin(*p), g.end(*p));
        out[u] = c;
    }
}
// Verify::compute_total_count (triangle_count/verifier.h:14-31): all directed pairs, sum / 6.
uint64_t tc_verify_total(const Graph &g) {
    uint64_t total = 0;
    #pragma omp parallel for schedule(dynamic, 64) reduction(+:total)
    for (int64_t u = 0; u < g.n; ++u)
        for (const vid *p = g.begin((vid)u); p != g.end((vid)u); ++p)
            total += isect_count(g.begin((vid)u), g.end((vid)u), g.begin(*p), g.end(*p));
    return total / 6;
}

// ---- k-cliques ------------------------------------------------------------------------------------------------------------
// Clique counting on a DAG: the quantity KcListing / Parallelize::{node,edge} return
// (k_clique_list/kernels/kclisting.h:163-188, parallelizationStrategy/parallelize.h:39-121): every k-clique
// of the underlying graph exactly once; k==1 -> #nodes, k==2 -> #directed edges.
uint64_t dag_cliques_rec(const Graph &g, int left, const vid *s, int64_t ns, std::vector<std::vector<vid>> &scratch) {
    // `left` more vertices to pick, all from s (candidates adjacent to everything picked so far)
    if (left == 1) return (uint64_t)ns;
    uint64_t c = 0;
    std::vector<vid> &buf = scratch[left];
    if ((int64_t)buf.size() < ns) buf.resize(ns);
    for (int64_t i = 0; i < ns; ++i) {
        vid v = s[i];
        if (left == 2) { c += isect_count(s, s + ns, g.begin(v), g.end(v)); continue; }
        int64_t k = isect_write(s, s + ns, g.begin(v), g.end(v), buf.data());
        if (k >= left - 1) c += dag_cliques_rec(g, left - 1, buf.data(), k, scratch);
    }
    return c;
}
// Algorithmic bytes of that recursion (B_k): every intersection S ∩ N+(v) it performs streams
// 4·(|S| + d+(v)) bytes.  Same walk as dag_cliques_rec with the byte counter added (the pruned branches have already
// paid for the intersection that pruned them).
uint64_t dag_clique_bytes_rec(const Graph &g, int left, const vid *s, int64_t ns, std::vector<std::vector<vid>> &scratch,
                              uint64_t &bytes) {
    if (left == 1) return (uint64_t)ns;
    uint64_t c = 0;
    std::vector<vid> &buf = scratch[left];
    if ((int64_t)buf.size() < ns) buf.resize(ns);
    for (int64_t i = 0; i < ns; ++i) {
        vid v = s[i];
        bytes += 4 * (uint64_t)(ns + g.deg(v));
        if (left == 2) { c += isect_count(s, s + ns, g.begin(v), g.end(v)); continue; }
        int64_t k = isect_write(s, s + ns, g.begin(v), g.end(v), buf.data());
        if (k >= left - 1) c += dag_clique_bytes_rec(g, left - 1, buf.data(), k, scratch, bytes);
    }
    return c;
}
uint64_t dag_clique_bytes(const Graph &g, int k, uint64_t *count_out) {
    uint64_t total = 0, bytes = 0;
    if (k >= 3) {
        #pragma omp parallel reduction(+:total, bytes)
        {
            std::vector<std::vector<vid>> scratch(k + 1);
            #pragma omp for schedule(dynamic, 16)
            for (int64_t u = 0; u < g.n; ++u)
                total += dag_clique_bytes_rec(g, k - 1, g.begin((vid)u), g.deg((vid)u), scratch, bytes);
        }
    }
    if (count_out) *count_out = k >= 3 ? total : (k == 2 ? (uint64_t)g.off[g.n] : (uint64_t)g.n);
    return bytes;
}
uint64_t dag_cliques(const Graph &g, int k) {
    if (k == 1) return (uint64_t)g.n;
    if (k == 2) return (uint64_t)g.off[g.n];
    uint64_t total = 0;
    #pragma omp parallel reduction(+:total)
    {
        std::vector<std::vector<vid>> scratch(k + 1);
        #pragma omp for schedule(dynamic, 16)
        for (int64_t u = 0; u < g.n; ++u)
            total += dag_cliques_rec(g, k - 1, g.begin((vid)u), g.deg((vid)u), scratch);
    }
    return total;
}
// Independent cross-check for larger k (NOT a reference restatement): clique counts for every size 1..kmax at once by
// pivoting on the undirected subgraph induced by N+(u) (succinct clique tree, Jain & Seshadhri, WSDM 2020).  It
// reproduces every reference-measured count of SURVEY.md §8c (e.g. kronecker-14: k=8 -> 138 220 170 775) with a
// different algorithm than both the reference and the CUDA kernels, so it pins k = 7..10 where the reference's
// one-by-one enumeration would take hours.  `dag` must be an oriented DAG (ids ascending along edges).
struct PivotSub {
    int D, W;
    std::vector<uint64_t> adj;          // D rows of W words, symmetric
};
unsigned __int128 binom128(int n, int r) {
    if (r < 0 || r > n) return 0;
    unsigned __int128 x = 1;
    for (int i = 0; i < r; ++i) x = x * (unsigned)(n - i) / (unsigned)(i + 1);
    return x;
}
void pivot_rec(const PivotSub &s, std::vector<uint64_t> &cand, int held, int pivots, int kmax, unsigned __int128 *cnt) {
    const int W = s.W;
    bool empty = true;
    for (int i = 0; i < W; ++i) if (cand[i]) { empty = false; break; }
    if (empty || held > kmax) {
        for (int k = held; k <= kmax && k <= held + pivots; ++k) cnt[k] += binom128(pivots, k - held);
        return;
    }
    int best = -1, best_c = -1;
    for (int w = 0; w < W; ++w)
        for (uint64_t m = cand[w]; m; m &= m - 1) {
            const int x = w * 64 + __builtin_ctzll(m);
            int c = 0;
            for (int i = 0; i < W; ++i) c += __builtin_popcountll(cand[i] & s.adj[(size_t)x * W + i]);
            if (c > best_c) { best_c = c; best = x; }
        }
    std::vector<uint64_t> nxt(W), todo(W);
    for (int i = 0; i < W; ++i) {
        nxt[i] = cand[i] & s.adj[(size_t)best * W + i];
        todo[i] = cand[i] & ~s.adj[(size_t)best * W + i];
    }
    todo[best / 64] &= ~(1ull << (best % 64));
    cand[best / 64] &= ~(1ull << (best % 64));
    pivot_rec(s, nxt, held, pivots + 1, kmax, cnt);              // pivot link: cliques may or may not use `best`
    for (int w = 0; w < W; ++w)
        for (uint64_t m = todo[w]; m; m &= m - 1) {
            const int b = __builtin_ctzll(m), x = w * 64 + b;
            for (int i = 0; i < W; ++i) nxt[i] = cand[i] & s.adj[(size_t)x * W + i];
            cand[w] &= ~(1ull << b);
            pivot_rec(s, nxt, held + 1, pivots, kmax, cnt);      // hold link: cliques that contain x
        }
}
void clique_counts_pivot(const Graph &dag, int kmax, uint64_t *out /* kmax+1 */) {
    std::vector<unsigned __int128> total(kmax + 1, 0);
    #pragma omp parallel
    {
        std::vector<unsigned __int128> local(kmax + 1, 0);
        #pragma omp for schedule(dynamic, 1)
        for (int64_t u = dag.n - 1; u >= 0; --u) {
            const vid *S = dag.begin((vid)u);
            const int D = (int)dag.deg((vid)u);
            PivotSub s;
            s.D = D; s.W = (D + 63) / 64;
            s.adj.assign((size_t)D * std::max(s.W, 1), 0);
            for (int i = 0; i < D; ++i)
                for (int j = i + 1; j < D; ++j)
                    if (std::binary_search(dag.begin(S[i]), dag.end(S[i]), S[j])) {
                        s.adj[(size_t)i * s.W + j / 64] |= 1ull << (j % 64);
                        s.adj[(size_t)j * s.W + i / 64] |= 1ull << (i % 64);
                    }
            std::vector<uint64_t> cand(std::max(s.W, 1), 0);
            for (int i = 0; i < D; ++i) cand[i / 64] |= 1ull << (i % 64);
            pivot_rec(s, cand, 1, 0, kmax, local.data());        // u itself is held
        }
        #pragma omp critical
        for (int k = 0; k <= kmax; ++k) total[k] += local[k];
    }
    for (int k = 0; k <= kmax; ++k) out[k] = (uint64_t)total[k];
}

// Set-based CliqueCount on the unoriented graph (k_clique_count/k_clique_count_set_based.h:6-31): counts
// ordered tuples, i.e. returns k!·C_k; the prune test |cur| >= k-2 uses the CURRENT k, kept verbatim.
uint64_t ordered_cliques_rec(const Graph &g, uint64_t k, const std::vector<vid> &s) {
    if (k == 1) return s.size();
    uint64_t c = 0;
    std::vector<vid> cur;
    for (vid v : s) {
        cur.resize(std::min<size_t>(s.size(), (size_t)g.deg(v)));
        cur.resize(isect_write(s.data(), s.data() + s.size(), g.begin(v), g.end(v), cur.data()));
        if (cur.size() >= k - 2) c += ordered_cliques_rec(g, k - 1, cur);
    }
    return c;
}
uint64_t ordered_cliques(const Graph &g, int k) {
    uint64_t total = 0;
    #pragma omp parallel for schedule(dynamic, 64) reduction(+:total)
    for (int64_t u = 0; u < g.n; ++u) {
        std::vector<vid> s(g.begin((vid)u), g.end((vid)u));
        total += ordered_cliques_rec(g, (uint64_t)k - 1, s);
    }
    return total;
}

// ---- vertex similarity (gms/algorithms/set_based/vertex_similarity/vertex_similarity.h) ---------------------------------------
// metric ids: 0 Jaccard 1 Overlap 2 AdamicAdar 3 Resource 4 CommNeigh 5 TotalNeigh 6 PrefAtt (enum order, :18)
double similarity(const Graph &g, int metric, vid a, vid b) {
    const vid *ab = g.begin(a), *ae = g.end(a), *bb = g.begin(b), *be = g.end(b);
    uint64_t da = ae - ab, db = be - bb;
    switch (metric) {
        case 0: {   // :30-37 — note the PLUS in the denominator and 1.0 for two empty sets
            if (da == 0 && db == 0) return 1.0;
            double c = (double)isect_count(ab, ae, bb, be);
            return c / (da + db + c);
        }
        case 1:     // :64-66 — 0/0 gives NaN as in the reference
            return double(isect_count(ab, ae, bb, be)) / std::min(da, db);
        case 2: case 3: {   // :95-106, :118-126 — ascending-id summation order
            std::vector<vid> w(std::min(da, db));
            int64_t k = isect_write(ab, ae, bb, be, w.data());
            double s = 0;
            for (int64_t i = 0; i < k; ++i) {
                double dw = (double)g.deg(w[i]);
                s += (metric == 2) ? 1. / std::log(dw) : 1.0 / dw;
            }
            return s;
        }
        case 4: return (double)isect_count(ab, ae, bb, be);          // :138-141
        case 5: return (double)union_count(ab, ae, bb, be);          // :153-156
        default: return (double)(da * db);                           // :168-170
    }
}

Graph *G(void *h) { return static_cast<Graph *>(h); }

}  // namespace

extern "C" {

void orc_set_threads(int t) { if (t > 0) omp_set_num_threads(t); }
int orc_max_threads() { return omp_get_max_threads(); }

void orc_generate_el(int scale, int degree, int uniform, int32_t *src, int32_t *dst) {
    int64_t m = (int64_t(1) << scale) * degree;
    if (uniform) uniform_edges(scale, m, src, dst);
    else rmat_edges(scale, m, 0.57f, 0.19f, 0.19f, true, src, dst);
}
void orc_rmat_el(int scale, int64_t m, float a, float b, float c, int permute, int32_t *src, int32_t *dst) {
    rmat_edges(scale, m, a, b, c, permute != 0, src, dst);
}
void *orc_from_el(int64_t m, const int32_t *src, const int32_t *dst, int symmetrize) {
    return new Graph(build_from_edges(m, src, dst, symmetrize != 0));
}
void *orc_generate(int scale, int degree, int uniform) {
    int64_t m = (int64_t(1) << scale) * degree;
    std::vector<vid> s(m), d(m);
    orc_generate_el(scale, degree, uniform, s.data(), d.data());
    return orc_from_el(m, s.data(), d.data(), 1);
}
void *orc_from_csr(int64_t n, const int64_t *off, const int32_t *nbr, int directed) {
    Graph *g = new Graph();
    g->n = n; g->directed = directed != 0;
    g->off.assign(off, off + n + 1);
    g->nbr.assign(nbr, nbr + off[n]);
    return g;
}
void orc_free(void *h) { delete G(h); }
int64_t orc_num_nodes(void *h) { return G(h)->n; }
int64_t orc_num_slots(void *h) { return G(h)->off[G(h)->n]; }
int orc_directed(void *h) { return G(h)->directed; }
void orc_export_csr(void *h, int64_t *off, int32_t *nbr) {
    std::copy(G(h)->off.begin(), G(h)->off.end(), off);
    std::copy(G(h)->nbr.begin(), G(h)->nbr.end(), nbr);
}
int orc_worth_relabelling(void *h) { return worth_relabelling(*G(h)); }
void *orc_relabel_by_degree(void *h) { return new Graph(relabel_by_degree(*G(h))); }

uint64_t orc_intersect_count(const int32_t *a, int64_t na, const int32_t *b, int64_t nb) {
    return isect_count(a, a + na, b, b + nb);
}
int64_t orc_intersect(const int32_t *a, int64_t na, const int32_t *b, int64_t nb, int32_t *out) {
    return isect_write(a, a + na, b, b + nb, out);
}
int64_t orc_union(const int32_t *a, int64_t na, const int32_t *b, int64_t nb, int32_t *out) {
    std::vector<vid> tmp(na + nb);
    int64_t k = std::set_union(a, a + na, b, b + nb, tmp.begin()) - tmp.begin();
    if (out) std::copy(tmp.begin(), tmp.begin() + k, out);
    return k;
}
uint64_t orc_union_count(const int32_t *a, int64_t na, const int32_t *b, int64_t nb) {
    return union_count(a, a + na, b, b + nb);
}
int64_t orc_difference(const int32_t *a, int64_t na, const int32_t *b, int64_t nb, int32_t *out) {
    return std::set_difference(a, a + na, b, b + nb, out) - out;
}
int orc_contains(const int32_t *a, int64_t na, int32_t x) { return std::binary_search(a, a + na, x); }

uint64_t orc_tc_total(void *h, int par) { return tc_total(*G(h), par != 0); }
double orc_tc_total_timed(void *h, int par, uint64_t *out) {
    auto t0 = std::chrono::steady_clock::now();
    *out = tc_total(*G(h), par != 0);
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
void orc_tc_vertex2(void *h, int /*variant*/, int64_t *out) { tc_vertex2(*G(h), out); }
uint64_t orc_tc_verify_total(void *h) { return tc_verify_total(*G(h)); }
double orc_tc_total_sample(void *h, int64_t stride, int64_t phase, int64_t *edges, uint64_t *sum) {
    const Graph &g = *G(h);
    std::vector<std::pair<vid, vid>> picks;
    int64_t idx = 0;
    for (int64_t u = 0; u < g.n; ++u)
        for (const vid *p = g.begin((vid)u); p != g.end((vid)u); ++p)
            if (u < *p) { if (idx % stride == phase) picks.emplace_back((vid)u, *p); ++idx; }
    uint64_t total = 0;
    int64_t np = (int64_t)picks.size();
    auto t0 = std::chrono::steady_clock::now();
    #pragma omp parallel for schedule(dynamic, 64) reduction(+:total)
    for (int64_t i = 0; i < np; ++i)
        total += isect_count(g.begin(picks[i].first), g.end(picks[i].first),
                             g.begin(picks[i].second), g.end(picks[i].second));
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    *edges = np; *sum = total;
    return dt;
}

void orc_degree_order(void *h, int rank_format, int32_t *out) { degree_order(*G(h), rank_format != 0, out); }
void orc_degeneracy_rank(void *h, int32_t *rank_out) { degeneracy_rank(*G(h), rank_out); }
int64_t orc_check_degeneracy_rank(void *h, const int32_t *rank) { return check_degeneracy_rank(*G(h), rank); }
// Approximate degeneracy order, averageDegree boundary (preprocessing/parallel/degeneracy_approx_csr.h:13-78,
// boundary_function.h:15-25): each round removes every remaining vertex with degree counter <= (unsigned)((1+eps) *
// mean counter of the remaining), ordered inside the round by that counter (the reference's parallel partition + sort
// leave ties unspecified; here ties go by id), then decrements the counter of EVERY neighbour (push style).
// round_of (optional) receives the round in which each vertex left, for batch-level comparison with the reference.
// boundary: 0 = averageDegree (boundary_function.h:15-25), 1 = minDegree (:27-35)
void orc_adg_order_ex(void *h, double eps, int rank_format, int boundary, int32_t *out, int32_t *round_of) {
    const Graph &g = *G(h);
    const int64_t n = g.n;
    std::vector<int> cnt(n);
    std::vector<vid> live(n);
    for (int64_t v = 0; v < n; ++v) { cnt[v] = (int)g.deg((vid)v); live[v] = (vid)v; }
    int64_t done = 0;
    int32_t round = 0;
    while (done < n) {
        unsigned int border;
        if (boundary == 0) {
            double res = 0;
            for (vid v : live) res += cnt[v];
            border = (unsigned int)((1 + eps) * (res / (double)live.size()));
        } else {
            int mn = 0x7fffffff;
            for (vid v : live) mn = std::min(mn, cnt[v]);
            border = (unsigned int)(2 * (1 + eps) * mn);
        }
        std::vector<vid> batch, rest;
        for (vid v : live) ((long long)cnt[v] <= (long long)border ? batch : rest).push_back(v);
        std::sort(batch.begin(), batch.end(), [&](vid a, vid b) { return cnt[a] < cnt[b] || (cnt[a] == cnt[b] && a < b); });
        for (size_t i = 0; i < batch.size(); ++i) {
            vid v = batch[i];
            if (rank_format) out[v] = (int32_t)(done + i); else out[done + i] = v;
            if (round_of) round_of[v] = round;
        }
        for (vid v : batch)
            for (const vid *p = g.begin(v); p != g.end(v); ++p) cnt[*p]--;
        done += (int64_t)batch.size();
        live.swap(rest);
        ++round;
    }
}

void orc_adg_order(void *h, double eps, int rank_format, int32_t *out, int32_t *round_of) {
    orc_adg_order_ex(h, eps, rank_format, 0, out, round_of);
}

// The reference's own acceptance test for a degeneracy order (verifiers/degeneracy_verifier.h:69-85), in rank
// format: with removal order = descending rank, every vertex may have at most `degeneracy` neighbours removed
// after it.  Returns max_v |{w in N(v): rank[w] < rank[v]}| (the "core number of the order"), or -1 if rank is
// not a permutation.  A valid order returns exactly the degeneracy.
int64_t orc_core_number_of_rank(void *h, const int32_t *rank) {
    const Graph &g = *G(h);
    std::vector<char> seen(g.n, 0);
    for (int64_t v = 0; v < g.n; ++v) {
        if (rank[v] < 0 || rank[v] >= g.n || seen[rank[v]]) return -1;
        seen[rank[v]] = 1;
    }
    int64_t worst = 0;
    for (int64_t v = 0; v < g.n; ++v) {
        int64_t later = 0;
        for (const vid *p = g.begin((vid)v); p != g.end((vid)v); ++p) later += rank[*p] < rank[v];
        worst = std::max(worst, later);
    }
    return worst;
}
void *orc_induce_directed(void *h, const int32_t *ranking) { return new Graph(induce_directed(*G(h), ranking)); }

uint64_t orc_kclique(void *h, int k, int /*mode*/) { return dag_cliques(*G(h), k); }
double orc_kclique_timed(void *h, int k, int mode, uint64_t *out) {
    auto t0 = std::chrono::steady_clock::now();
    *out = orc_kclique(h, k, mode);
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
uint64_t orc_clique_count_set_based(void *h, int k) { return ordered_cliques(*G(h), k); }
uint64_t orc_kclique_bytes(void *h, int k, uint64_t *count_out) { return dag_clique_bytes(*G(h), k, count_out); }
void orc_clique_counts_pivot(void *h, int kmax, uint64_t *out) { clique_counts_pivot(*G(h), kmax, out); }

double orc_vertex_similarity(void *h, int metric, int32_t a, int32_t b) { return similarity(*G(h), metric, a, b); }
void orc_pair_similarity(void *h, int metric, int64_t np, const int32_t *a, const int32_t *b, double *out) {
    const Graph &g = *G(h);
    #pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < np; ++i) out[i] = similarity(g, metric, a[i], b[i]);
}
int64_t orc_edge_similarity(void *h, int metric, double *out) {
    const Graph &g = *G(h);
    std::vector<int64_t> base(g.n + 1, 0);
    for (int64_t u = 0; u < g.n; ++u) {
        int64_t c = 0;
        for (const vid *p = g.begin((vid)u); p != g.end((vid)u); ++p) c += (u < *p);
        base[u + 1] = base[u] + c;
    }
    if (out) {
        #pragma omp parallel for schedule(dynamic, 64)
        for (int64_t u = 0; u < g.n; ++u) {
            int64_t pos = base[u];
            for (const vid *p = g.begin((vid)u); p != g.end((vid)u); ++p)
                if (u < *p) out[pos++] = similarity(g, metric, (vid)u, *p);
        }
    }
    return base[g.n];
}

// ---- bookkeeping helpers for tests (not reference behaviour) --------------------------------------------------------------------
// B_TC = Σ_{(u,v)∈E⁺} 4·(d⁺(u)+d⁺(v)) on the (degree asc, id asc)-oriented DAG, and
// B_ref = Σ_{u<v} 4·(d(u)+d(v)) — what the reference's unoriented loop streams.
void orc_tc_bytes(void *h, uint64_t *b_tc, uint64_t *b_ref, int64_t *max_dplus) {
    const Graph &g = *G(h);
    std::vector<vid> rank(g.n);
    degree_order(g, true, rank.data());
    std::vector<int64_t> dplus(g.n, 0);
    for (int64_t u = 0; u < g.n; ++u)
        for (const vid *p = g.begin((vid)u); p != g.end((vid)u); ++p) dplus[u] += rank[u] < rank[*p];
    uint64_t bt = 0, br = 0; int64_t mx = 0;
    for (int64_t u = 0; u < g.n; ++u) {
        mx = std::max(mx, dplus[u]);
        for (const vid *p = g.begin((vid)u); p != g.end((vid)u); ++p) {
            if (rank[u] < rank[*p]) bt += 4 * uint64_t(dplus[u] + dplus[*p]);
            if (u < *p) br += 4 * uint64_t(g.deg((vid)u) + g.deg(*p));
        }
    }
    *b_tc = bt; *b_ref = br; *max_dplus = mx;
}

}  // extern "C"
