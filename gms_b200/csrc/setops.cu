// setops.cu — batched SortedSet algebra over neighbourhoods, and the vertex-similarity measures built on it.
//
// Replaces:
//   SortedSetBase::intersect_count / intersect      gms/representations/sets/sorted_set.h:160-182
//   SortedSetBase::union_with / union_count / difference   sorted_set.h:104-109,140,184-189
//   GMS::VertexSim::vertex_similarity<Metric>       gms/algorithms/set_based/vertex_similarity/vertex_similarity.h:30-221
//
// One warp per vertex pair; merge path for balanced pairs, galloping for skewed ones (isect.cuh).  The similarity
// scores are IEEE doubles computed from integer counts with the same expressions as the reference, so Jaccard,
// Overlap, CommNeigh, TotalNeigh and PrefAtt are bit-exact; Resource sums 1/deg(w) in ascending-w order like the
// reference (bit-exact); AdamicAdar additionally depends on log() and agrees to a few ulp.
#include "common.cuh"
#include "sort.cuh"
#include "isect.cuh"
#include "ops.cuh"

namespace gmsb {

namespace {

constexpr int kWarps = 8;
constexpr int kGallopRatio = 8;

__device__ __forceinline__ uint32_t pair_count(const vid_t *a, int na, const vid_t *b, int nb, int lane, vid_t *buf) {
    int lo = na < nb ? na : nb, hi = na < nb ? nb : na;
    if (lo == 0) return 0;
    if ((long long)hi >= (long long)kGallopRatio * lo) return warp_gallop_count(a, na, b, nb, lane);
    return warp_merge_count(a, na, b, nb, lane, buf);
}

__global__ void __launch_bounds__(kWarps * 32)
k_pair_count(const eid_t *__restrict__ off, const vid_t *__restrict__ nbr, int64_t n, int64_t np,
             const vid_t *__restrict__ pa, const vid_t *__restrict__ pb, unsigned long long *__restrict__ out,
             int *__restrict__ bad, int union_mode) {
    __shared__ vid_t stage[kWarps][kMergeTile + 2];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < np; i += nwarps) {
        const vid_t a = pa[i], b = pb[i];
        if (a < 0 || b < 0 || a >= n || b >= n) { if (lane == 0) { *bad = 1; out[i] = 0; } continue; }
        const eid_t oa = off[a], ob = off[b];
        unsigned long long c = pair_count(nbr + oa, (int)(off[a + 1] - oa), nbr + ob, (int)(off[b + 1] - ob), lane,
                                          stage[wib]);
        c = warp_sum(c);
        // union_count = |A| + |B| - |A ∩ B|   (sorted_set.h:140)
        if (lane == 0) out[i] = union_mode ? (unsigned long long)(off[a + 1] - oa) + (unsigned long long)(off[b + 1] - ob) - c : c;
    }
}

// Materialising intersection, ascending order: lanes walk the shorter list in chunks of 32, ballots keep the order.
// pass 0 (out_elems == nullptr) only counts; pass 1 writes at out_off[i].
__global__ void __launch_bounds__(256)
k_pair_intersect(const eid_t *__restrict__ off, const vid_t *__restrict__ nbr, int64_t np, const vid_t *__restrict__ pa,
                 const vid_t *__restrict__ pb, int64_t *__restrict__ counts, const int64_t *__restrict__ out_off,
                 vid_t *__restrict__ out_elems) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < np; i += nwarps) {
        const vid_t va = pa[i], vb = pb[i];
        const vid_t *a = nbr + off[va], *b = nbr + off[vb];
        int na = (int)(off[va + 1] - off[va]), nb = (int)(off[vb + 1] - off[vb]);
        if (na > nb) { const vid_t *t = a; a = b; b = t; int tn = na; na = nb; nb = tn; }
        int64_t w = out_elems ? out_off[i] : 0;
        int64_t c = 0;
        for (int j0 = 0; j0 < na; j0 += 32) {
            const int j = j0 + lane;
            bool hit = false;
            vid_t x = 0;
            if (j < na) {
                x = a[j];
                int lo = lower_bound_dev(b, nb, x);
                hit = lo < nb && b[lo] == x;
            }
            unsigned mask = __ballot_sync(0xffffffffu, hit);
            if (out_elems && hit) out_elems[w + c + __popc(mask & ((1u << lane) - 1))] = x;
            c += __popc(mask);
        }
        if (!out_elems && lane == 0) counts[i] = c;
    }
}

// Materialising difference N(a) \\ N(b), ascending (std::set_difference, sorted_set_operations.h:74-79): lanes walk A in
// chunks of 32 and keep the elements that are NOT in B; same two-pass protocol as k_pair_intersect.
__global__ void __launch_bounds__(256)
k_pair_difference(const eid_t *__restrict__ off, const vid_t *__restrict__ nbr, int64_t np, const vid_t *__restrict__ pa,
                  const vid_t *__restrict__ pb, int64_t *__restrict__ counts, const int64_t *__restrict__ out_off,
                  vid_t *__restrict__ out_elems) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < np; i += nwarps) {
        const vid_t va = pa[i], vb = pb[i];
        const vid_t *a = nbr + off[va], *b = nbr + off[vb];
        const int na = (int)(off[va + 1] - off[va]), nb = (int)(off[vb + 1] - off[vb]);
        const int64_t w = out_elems ? out_off[i] : 0;
        int64_t c = 0;
        for (int j0 = 0; j0 < na; j0 += 32) {
            const int j = j0 + lane;
            bool keep = false;
            vid_t x = 0;
            if (j < na) {
                x = a[j];
                const int lo = lower_bound_dev(b, nb, x);
                keep = !(lo < nb && b[lo] == x);
            }
            const unsigned mask = __ballot_sync(0xffffffffu, keep);
            if (out_elems && keep) out_elems[w + c + __popc(mask & ((1u << lane) - 1))] = x;
            c += __popc(mask);
        }
        if (!out_elems && lane == 0) counts[i] = c;
    }
}

// Materialising union, ascending (std::set_union, sorted_set_operations.h:30-35).  Every element knows its place in
// the output without a merge: an element x of A sits after the |A<x| + |B<x| - |A∩B<x| smaller ones, i.e. at
// j + lower_bound(B, x) - (common elements among A[0..j)); an element of B that is not in A likewise.
__global__ void __launch_bounds__(256)
k_pair_union(const eid_t *__restrict__ off, const vid_t *__restrict__ nbr, int64_t np, const vid_t *__restrict__ pa,
             const vid_t *__restrict__ pb, int64_t *__restrict__ counts, const int64_t *__restrict__ out_off,
             vid_t *__restrict__ out_elems) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < np; i += nwarps) {
        const vid_t va = pa[i], vb = pb[i];
        const vid_t *a = nbr + off[va], *b = nbr + off[vb];
        const int na = (int)(off[va + 1] - off[va]), nb = (int)(off[vb + 1] - off[vb]);
        const int64_t w = out_elems ? out_off[i] : 0;
        int common = 0;                      // common elements seen so far (running over the chunks)
        for (int j0 = 0; j0 < na; j0 += 32) {
            const int j = j0 + lane;
            bool hit = false;
            int lo = 0;
            vid_t x = 0;
            if (j < na) {
                x = a[j];
                lo = lower_bound_dev(b, nb, x);
                hit = lo < nb && b[lo] == x;
            }
            const unsigned mask = __ballot_sync(0xffffffffu, hit);
            if (out_elems && j < na) out_elems[w + j + lo - (common + __popc(mask & ((1u << lane) - 1)))] = x;
            common += __popc(mask);
        }
        if (!out_elems) {
            if (lane == 0) counts[i] = (int64_t)na + nb - common;
            continue;
        }
        common = 0;
        for (int t0 = 0; t0 < nb; t0 += 32) {
            const int t = t0 + lane;
            bool hit = false;
            int lo = 0;
            vid_t x = 0;
            if (t < nb) {
                x = b[t];
                lo = lower_bound_dev(a, na, x);
                hit = lo < na && a[lo] == x;
            }
            const unsigned mask = __ballot_sync(0xffffffffu, hit);
            if (t < nb && !hit) out_elems[w + t + lo - (common + __popc(mask & ((1u << lane) - 1)))] = x;
            common += __popc(mask);
        }
    }
}

// Similarity of one pair; lane 0 writes the score.
__global__ void __launch_bounds__(kWarps * 32)
k_pair_similarity(const eid_t *__restrict__ off, const vid_t *__restrict__ nbr, int64_t n, int metric, int64_t np,
                  const vid_t *__restrict__ pa, const vid_t *__restrict__ pb, double *__restrict__ out,
                  int *__restrict__ bad) {
    __shared__ vid_t stage[kWarps][kMergeTile + 2];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < np; i += nwarps) {
        const vid_t va = pa[i], vb = pb[i];
        if (va < 0 || vb < 0 || va >= n || vb >= n) { if (lane == 0) { *bad = 1; out[i] = 0; } continue; }
        const vid_t *a = nbr + off[va], *b = nbr + off[vb];
        const int na = (int)(off[va + 1] - off[va]), nb = (int)(off[vb + 1] - off[vb]);
        double score;
        if (metric == GMSB_SIM_PREF_ATT) {
            score = (double)((unsigned long long)na * (unsigned long long)nb);            // vertex_similarity.h:153-156
        } else if (metric == GMSB_SIM_ADAMIC_ADAR || metric == GMSB_SIM_RESOURCE) {
            // sum over w in A∩B, ascending w, of 1/log(deg w) or 1/deg w  (vertex_similarity.h:95-126)
            const vid_t *s = a, *l = b;
            int ns = na, nl = nb;
            if (ns > nl) { const vid_t *t = s; s = l; l = t; int tn = ns; ns = nl; nl = tn; }
            double sum = 0;
            for (int j0 = 0; j0 < ns; j0 += 32) {
                const int j = j0 + lane;
                bool hit = false;
                double term = 0;
                if (j < ns) {
                    const vid_t x = s[j];
                    int lo = lower_bound_dev(l, nl, x);
                    hit = lo < nl && l[lo] == x;
                    if (hit) {
                        const double dw = (double)(off[x + 1] - off[x]);
                        term = metric == GMSB_SIM_ADAMIC_ADAR ? 1. / log(dw) : 1.0 / dw;
                    }
                }
                unsigned mask = __ballot_sync(0xffffffffu, hit);
                while (mask) {                         // sequential, ascending: same rounding order as the reference
                    const int src = __ffs(mask) - 1;
                    sum += __shfl_sync(0xffffffffu, term, src);
                    mask &= mask - 1;
                }
            }
            score = sum;
        } else {
            unsigned long long c = warp_sum(pair_count(a, na, b, nb, lane, stage[wib]));
            const double cd = (double)c;
            if (metric == GMSB_SIM_JACCARD)                                              // :30-37 (sic: plus)
                score = (na == 0 && nb == 0) ? 1.0 : cd / (double)((unsigned long long)na + (unsigned long long)nb + cd);
            else if (metric == GMSB_SIM_OVERLAP)                                         // :64-66
                score = cd / (double)(unsigned long long)(na < nb ? na : nb);
            else if (metric == GMSB_SIM_COMM_NEIGH)                                      // :138-141
                score = cd;
            else                                                                         // TotalNeigh = |A ∪ B|
                score = (double)((unsigned long long)na + (unsigned long long)nb - c);
        }
        if (lane == 0) out[i] = score;
    }
}

// undirected edges u<v in CSR order
__global__ void k_count_upper(const eid_t *__restrict__ off, const vid_t *__restrict__ nbr, int64_t n,
                              int64_t *__restrict__ cnt) {
    for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < n; u += (int64_t)gridDim.x * blockDim.x) {
        const eid_t b = off[u], e = off[u + 1];
        // lists are ascending: entries > u form a suffix
        int lo = lower_bound_dev(nbr + b, (int)(e - b), (vid_t)(u + 1));
        cnt[u] = (e - b) - lo;
        if (u == 0) cnt[n] = 0;
    }
}
__global__ void k_emit_upper(const eid_t *__restrict__ off, const vid_t *__restrict__ nbr, int64_t n,
                             const int64_t *__restrict__ base, vid_t *__restrict__ pa, vid_t *__restrict__ pb) {
    int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warp; u < n; u += nwarps) {
        const eid_t e = off[u + 1];
        const int64_t c = base[u + 1] - base[u];
        const eid_t s0 = e - c;
        for (int64_t j = lane; j < c; j += 32) { pa[base[u] + j] = (vid_t)u; pb[base[u] + j] = nbr[s0 + j]; }
    }
}

void check_bad(DevBuf<int> &bad, const char *what) {
    GMSB_REQUIRE(bad.get(0) == 0, std::string(what) + ": vertex id out of range");
}

}  // namespace

static void count_batch(Graph &g, int64_t np, const vid_t *a, const vid_t *b, uint64_t *out, bool union_mode) {
    if (np == 0) return;
    Runtime &r = rt();
    DevBuf<vid_t> da(np), db(np);
    DevBuf<unsigned long long> dout(np);
    DevBuf<int> bad(1);
    bad.zero();
    da.upload(a, np); db.upload(b, np);
    int grid = (int)std::min<int64_t>(ceil_div(np, kWarps), (int64_t)r.sm_count * 16);
    k_pair_count<<<grid, kWarps * 32, 0, r.stream>>>(g.off.p, g.nbr.p, g.n, np, da.p, db.p, dout.p, bad.p, union_mode ? 1 : 0);
    launched();
    dout.download(reinterpret_cast<unsigned long long *>(out), np);
    check_bad(bad, union_mode ? "union_count_batch" : "intersect_count_batch");
}
void intersect_count_batch(Graph &g, int64_t np, const vid_t *a, const vid_t *b, uint64_t *out) {
    count_batch(g, np, a, b, out, false);
}
void union_count_batch(Graph &g, int64_t np, const vid_t *a, const vid_t *b, uint64_t *out) {
    count_batch(g, np, a, b, out, true);
}

namespace {
enum class SetOp { Intersect, Difference, Union };
void set_batch(Graph &g, SetOp op, const char *what, int64_t np, const vid_t *a, const vid_t *b, int64_t *out_offsets,
               vid_t *out_elems, int64_t cap) {
    out_offsets[0] = 0;
    if (np == 0) return;
    Runtime &r = rt();
    for (int64_t i = 0; i < np; ++i)
        GMSB_REQUIRE(a[i] >= 0 && a[i] < g.n && b[i] >= 0 && b[i] < g.n, std::string(what) + ": vertex id out of range");
    DevBuf<vid_t> da(np), db(np);
    DevBuf<int64_t> cnt(np + 1), pos(np + 1);
    cnt.zero();
    da.upload(a, np); db.upload(b, np);
    const int grid = (int)std::min<int64_t>(ceil_div(np, 8), (int64_t)r.sm_count * 16);
    auto launch = [&](int64_t *counts, const int64_t *offs, vid_t *elems) {
        if (op == SetOp::Intersect)
            k_pair_intersect<<<grid, 256, 0, r.stream>>>(g.off.p, g.nbr.p, np, da.p, db.p, counts, offs, elems);
        else if (op == SetOp::Difference)
            k_pair_difference<<<grid, 256, 0, r.stream>>>(g.off.p, g.nbr.p, np, da.p, db.p, counts, offs, elems);
        else
            k_pair_union<<<grid, 256, 0, r.stream>>>(g.off.p, g.nbr.p, np, da.p, db.p, counts, offs, elems);
        launched();
    };
    launch(cnt.p, nullptr, nullptr);
    exclusive_sum(cnt.p, pos.p, np + 1);
    pos.download(out_offsets, np + 1);
    const int64_t total = out_offsets[np];
    if (!out_elems) return;
    GMSB_REQUIRE(cap >= total, std::string(what) + ": output capacity too small");
    if (total == 0) return;
    DevBuf<vid_t> elems(total);
    launch(nullptr, pos.p, elems.p);
    elems.download(out_elems, total);
}
}  // namespace

void intersect_batch(Graph &g, int64_t np, const vid_t *a, const vid_t *b, int64_t *out_offsets, vid_t *out_elems,
                     int64_t cap) {
    set_batch(g, SetOp::Intersect, "intersect_batch", np, a, b, out_offsets, out_elems, cap);
}
void difference_batch(Graph &g, int64_t np, const vid_t *a, const vid_t *b, int64_t *out_offsets, vid_t *out_elems,
                      int64_t cap) {
    set_batch(g, SetOp::Difference, "difference_batch", np, a, b, out_offsets, out_elems, cap);
}
void union_batch(Graph &g, int64_t np, const vid_t *a, const vid_t *b, int64_t *out_offsets, vid_t *out_elems,
                 int64_t cap) {
    set_batch(g, SetOp::Union, "union_batch", np, a, b, out_offsets, out_elems, cap);
}

void pair_similarity_device(Graph &g, int metric, int64_t np, const vid_t *da, const vid_t *db, double *out) {
    Runtime &r = rt();
    DevBuf<double> dout(np);
    DevBuf<int> bad(1);
    bad.zero();
    int grid = (int)std::min<int64_t>(ceil_div(np, kWarps), (int64_t)r.sm_count * 16);
    k_pair_similarity<<<grid, kWarps * 32, 0, r.stream>>>(g.off.p, g.nbr.p, g.n, metric, np, da, db, dout.p, bad.p);
    launched();
    dout.download(out, np);
    check_bad(bad, "pair_similarity");
}

void pair_similarity(Graph &g, int metric, int64_t np, const vid_t *a, const vid_t *b, double *out) {
    if (np == 0) return;
    DevBuf<vid_t> da(np), db(np);
    da.upload(a, np); db.upload(b, np);
    pair_similarity_device(g, metric, np, da.p, db.p, out);
}

void upper_edge_base(Graph &g, DevBuf<int64_t> &base, int64_t *m_out) {
    Runtime &r = rt();
    const int64_t n = g.n;
    DevBuf<int64_t> cnt(n + 1);
    base.alloc(n + 1);
    *m_out = 0;
    if (n) {
        k_count_upper<<<grid_for(n, 256), 256, 0, r.stream>>>(g.off.p, g.nbr.p, n, cnt.p); launched();
        exclusive_sum(cnt.p, base.p, n + 1);
        *m_out = base.get(n);
    }
}
void emit_upper_pairs(Graph &g, const int64_t *base_dev, vid_t *pa, vid_t *pb) {
    k_emit_upper<<<grid_for(g.n * 32, 256), 256, 0, rt().stream>>>(g.off.p, g.nbr.p, g.n, base_dev, pa, pb); launched();
}

void edge_similarity(Graph &g, int metric, double *out, int64_t *m_out) {
    GMSB_REQUIRE(!g.directed, "edge_similarity: graph must be undirected");
    Runtime &r = rt();
    const int64_t n = g.n;
    int64_t m = 0;
    DevBuf<int64_t> base;
    upper_edge_base(g, base, &m);
    if (m_out) *m_out = m;
    if (!out || m == 0) return;
    if (metric != GMSB_SIM_ADAMIC_ADAR && metric != GMSB_SIM_RESOURCE) {
        // these five depend on the graph only through |N(a) ∩ N(b)| = the edge's triangle support: one pass of the
        // oriented triangle schedule instead of one symmetric intersection per edge (tc_support.cu)
        DevBuf<double> dout(m);
        edge_scores_from_support(g, metric, base.p, dout.p);
        dout.download(out, m);
        return;
    }
    DevBuf<vid_t> pa(m), pb(m);
    k_emit_upper<<<grid_for(n * 32, 256), 256, 0, r.stream>>>(g.off.p, g.nbr.p, n, base.p, pa.p, pb.p); launched();
    pair_similarity_device(g, metric, m, pa.p, pb.p, out);
}

}  // namespace gmsb
