// orient.cu — vertex orderings and DAG orientation on the device.
//
// Replaces:
//   PpParallel::getDegreeOrdering            gms/algorithms/preprocessing/parallel/degree.h:16-61
//   PpSequential::InduceDirectedGraph        gms/algorithms/preprocessing/sequential/apply_order.h:10-35
//   PpSequential::getDegeneracyOrderingDanischHeap  gms/algorithms/preprocessing/sequential/degeneracy_danisch.h:12-56
//
// The oriented graph lives in RANK SPACE: vertex ids are replaced by their position in the ordering, so
// N+(u) = { w in N(u) : w > u } and every list is a suffix-closed, ascending id range.  That makes "v is a hub"
// equivalent to "v is close to n", which is what lets the bitmap kernel (tc.cu) keep N+(v) in shared memory.
#include "common.cuh"
#include "sort.cuh"
#include "isect.cuh"
#include "orient.cuh"

#include <cstdlib>

namespace gmsb {

namespace {


__global__ void k_degree_keys(const eid_t *__restrict__ off, int64_t n, uint64_t *__restrict__ keys) {
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x)
        keys[v] = ((uint64_t)(off[v + 1] - off[v]) << 32) | (uint32_t)v;
}

__global__ void k_order_rank(const uint64_t *__restrict__ sorted, int64_t n, vid_t *__restrict__ order,
                             vid_t *__restrict__ rank) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        vid_t v = (vid_t)(uint32_t)sorted[i];
        order[i] = v;
        rank[v] = (vid_t)i;
    }
}

// One warp per vertex: d+(u) = |{ v in N(u) : rank[v] > rank[u] }|, stored at cnt[rank[u] + 1].
__global__ void k_count_out(const eid_t *__restrict__ off, const vid_t *__restrict__ nbr, int64_t n,
                            const vid_t *__restrict__ rank, eid_t *__restrict__ cnt) {
    int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warp; u < n; u += nwarps) {
        eid_t b = off[u], e = off[u + 1];
        vid_t ru = rank[u];
        int c = 0;
        for (eid_t s = b + lane; s < e; s += 32) c += rank[nbr[s]] > ru;
        for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) cnt[ru + 1] = c;
    }
}

// One warp per vertex: write (rank[u]<<32 | rank[v]) for the kept slots, compacted with ballots, into the
// segment of rank[u]; the order inside a segment is fixed afterwards by the radix sort.
__global__ void k_emit_out(const eid_t *__restrict__ off, const vid_t *__restrict__ nbr, int64_t n,
                           const vid_t *__restrict__ rank, const eid_t *__restrict__ doff,
                           uint64_t *__restrict__ keys) {
    int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warp; u < n; u += nwarps) {
        eid_t b = off[u], e = off[u + 1];
        vid_t ru = rank[u];
        eid_t w = doff[ru];
        uint64_t hi = (uint64_t)(uint32_t)ru << 32;
        for (eid_t s0 = b; s0 < e; s0 += 32) {
            eid_t s = s0 + lane;
            vid_t rv = s < e ? rank[nbr[s]] : -1;
            bool kp = rv > ru;
            unsigned mask = __ballot_sync(0xffffffffu, kp);
            if (kp) keys[w + __popc(mask & ((1u << lane) - 1))] = hi | (uint32_t)rv;
            w += __popc(mask);
        }
    }
}

__global__ void k_low32(const uint64_t *__restrict__ keys, int64_t K, vid_t *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < K; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (vid_t)(uint32_t)keys[i];
}

__global__ void k_max_deg(const eid_t *__restrict__ off, int64_t n, int *out) {
    int mx = 0;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x)
        mx = max(mx, (int)(off[v + 1] - off[v]));
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, mx);
}

// last vertex with a non-empty list, +1 (InduceDirectedGraph's n = max id seen in the edge list + 1)
__global__ void k_max_endpoint(const eid_t *__restrict__ off, const vid_t *__restrict__ nbr, int64_t n, int *out) {
    int mx = 0;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        eid_t b = off[v], e = off[v + 1];
        if (e > b) mx = max(mx, max((int)v, (int)nbr[e - 1]));
    }
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, mx);
}

__global__ void k_check_perm(const vid_t *__restrict__ rank, int64_t n, int *__restrict__ seen, int *bad) {
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        vid_t r = rank[v];
        if (r < 0 || r >= n || atomicExch(&seen[r], 1)) *bad = 1;
    }
}

// ---- fast path: per-list sort on chip instead of a global 64-bit radix sort -----------------------------------
// (1) relabel + count: rnbr[s] = rank[nbr[s]] is stored so that the emit pass streams instead of gathering again.
__global__ void k_relabel_count(const eid_t *__restrict__ off, const vid_t *__restrict__ nbr, int64_t n,
                                const vid_t *__restrict__ rank, vid_t *__restrict__ rnbr, eid_t *__restrict__ cnt) {
    int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warp; u < n; u += nwarps) {
        eid_t b = off[u], e = off[u + 1];
        vid_t ru = rank[u];
        int c = 0;
        for (eid_t s = b + lane; s < e; s += 32) {
            vid_t rv = rank[nbr[s]];
            rnbr[s] = rv;
            c += rv > ru;
        }
        for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) cnt[ru + 1] = c;
    }
}

__device__ __forceinline__ vid_t warp_bitonic_sort(vid_t v, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const vid_t other = __shfl_xor_sync(0xffffffffu, v, j);
            const bool up = (lane & k) == 0, lower = (lane & j) == 0;
            v = (lower == up) ? min(v, other) : max(v, other);
        }
    }
    return v;
}

// (2) emit: one warp per vertex compacts the higher-ranked neighbours with ballots.  Lists of <= 32 survivors are
// sorted in registers (shuffle bitonic network) and written in final order; longer ones are written unsorted and
// queued for the CTA sorter.
__global__ void __launch_bounds__(256)
k_emit_sorted(const eid_t *__restrict__ off, const vid_t *__restrict__ rnbr, int64_t n, const vid_t *__restrict__ rank,
              const eid_t *__restrict__ doff, vid_t *__restrict__ dnbr, vid_t *__restrict__ big, int *__restrict__ nbig) {
    __shared__ vid_t stage[8][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warp; u < n; u += nwarps) {
        const vid_t ru = rank[u];
        const eid_t ob = doff[ru];
        const int c = (int)(doff[ru + 1] - ob);
        if (c == 0) continue;
        const eid_t b = off[u], e = off[u + 1];
        const bool small = c <= 32;
        int w = 0;
        __syncwarp();
        for (eid_t s0 = b; s0 < e; s0 += 32) {
            const eid_t s = s0 + lane;
            const vid_t rv = s < e ? rnbr[s] : -1;
            const bool kp = rv > ru;
            const unsigned mask = __ballot_sync(0xffffffffu, kp);
            if (kp) {
                const int p = w + __popc(mask & ((1u << lane) - 1));
                if (small) stage[wib][p] = rv; else dnbr[ob + p] = rv;
            }
            w += __popc(mask);
        }
        if (small) {
            __syncwarp();
            vid_t v = lane < c ? stage[wib][lane] : 0x7fffffff;
            v = warp_bitonic_sort(v, lane);
            if (lane < c) dnbr[ob + lane] = v;
        } else if (lane == 0) {
            // two queues filled from opposite ends of one array: [0, nbig[0]) mid-size lists, (n-1-nbig[1], n-1] long
            if (c <= 512) big[atomicAdd(&nbig[0], 1)] = ru;
            else big[n - 1 - atomicAdd(&nbig[1], 1)] = ru;
        }
    }
}

// (3a) warp sorter for queued lists of 33..kWarpSortCap elements: one warp per list, bitonic network in the warp's
// own shared-memory slice (no block barriers), eight lists per CTA in flight.
constexpr int kWarpSortCap = 512;
__global__ void __launch_bounds__(256)
k_sort_mid(const vid_t *__restrict__ big, int nbig, const eid_t *__restrict__ doff, vid_t *__restrict__ dnbr) {
    __shared__ vid_t bufs[8][kWarpSortCap];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    vid_t *buf = bufs[wib];
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < nbig; t += nwarps) {
        const vid_t ru = big[t];
        const eid_t ob = doff[ru];
        const int c = (int)(doff[ru + 1] - ob);
        if (c > kWarpSortCap) continue;                    // left to the CTA sorter
        int P = 64;
        while (P < c) P <<= 1;
        __syncwarp();
        for (int i = lane; i < P; i += 32) buf[i] = i < c ? dnbr[ob + i] : 0x7fffffff;
        __syncwarp();
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int q = lane; q < (P >> 1); q += 32) {
                    // q-th compare-exchange pair of this step: insert a zero bit at position log2(j)
                    const int i = ((q & ~(j - 1)) << 1) | (q & (j - 1));
                    const int x = i | j;
                    const vid_t a = buf[i], bq = buf[x];
                    const bool up = (i & k) == 0;
                    if ((a > bq) == up) { buf[i] = bq; buf[x] = a; }
                }
                __syncwarp();
            }
        }
        for (int i = lane; i < c; i += 32) dnbr[ob + i] = buf[i];
    }
}

// (3b) CTA sorter for the few longer lists: bitonic network in shared memory, sized per list.
constexpr int kSortCap = 8192;
__global__ void __launch_bounds__(256)
k_sort_big(const vid_t *__restrict__ big, int nbig, const eid_t *__restrict__ doff, vid_t *__restrict__ dnbr) {
    extern __shared__ vid_t buf[];
    for (int t = blockIdx.x; t < nbig; t += gridDim.x) {
        const vid_t ru = big[t];
        const eid_t ob = doff[ru];
        const int c = (int)(doff[ru + 1] - ob);
        if (c <= kWarpSortCap) continue;                   // done by the warp sorter
        int P = 64;
        while (P < c) P <<= 1;
        __syncthreads();
        for (int i = threadIdx.x; i < P; i += blockDim.x) buf[i] = i < c ? dnbr[ob + i] : 0x7fffffff;
        __syncthreads();
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < P; i += blockDim.x) {
                    const int x = i ^ j;
                    if (x > i) {
                        const vid_t a = buf[i], bq = buf[x];
                        const bool up = (i & k) == 0;
                        if ((a > bq) == up) { buf[i] = bq; buf[x] = a; }
                    }
                }
                __syncthreads();
            }
        }
        for (int i = threadIdx.x; i < c; i += blockDim.x) dnbr[ob + i] = buf[i];
    }
}

}  // namespace

void degree_order(const Graph &g, DevBuf<vid_t> &order, DevBuf<vid_t> &rank) {
    Runtime &r = rt();
    int64_t n = g.n;
    order.alloc(n);
    rank.alloc(n);
    if (n == 0) return;
    DevBuf<uint64_t> keys(n), alt(n);
    k_degree_keys<<<grid_for(n, 256), 256, 0, r.stream>>>(g.off.p, n, keys.p); launched();
    uint64_t *sorted = radix_sort_keys(keys.p, alt.p, n, 0, 64);     // (degree asc, id asc)
    k_order_rank<<<grid_for(n, 256), 256, 0, r.stream>>>(sorted, n, order.p, rank.p); launched();
}

void orient_by_rank(const Graph &g, const vid_t *rank_dev, DevBuf<eid_t> &doff, DevBuf<vid_t> &dnbr, int64_t *m_out,
                    int *max_dplus) {
    Runtime &r = rt();
    int64_t n = g.n;
    doff.alloc(n + 1);
    doff.zero();
    DevBuf<vid_t> rnbr(g.slots);
    if (n) {
        k_relabel_count<<<grid_for(n * 32, 256), 256, 0, r.stream>>>(g.off.p, g.nbr.p, n, rank_dev, rnbr.p, doff.p);
        launched();
        inclusive_sum_inplace(doff.p, n + 1);
    }
    int64_t m = n ? doff.get(n) : 0;
    *m_out = m;
    dnbr.alloc(m);
    int maxd = 0;
    if (n) {
        DevBuf<int> mx(1);
        mx.zero();
        k_max_deg<<<grid_for(n, 256), 256, 0, r.stream>>>(doff.p, n, mx.p); launched();
        maxd = mx.get(0);
    }
    if (max_dplus) *max_dplus = maxd;
    if (m == 0) return;
    int cap = kSortCap;
    if (const char *env = getenv("GMSB_ORIENT_SORT_CAP")) cap = std::min(kSortCap, std::max(0, atoi(env)));   // tests
    if (maxd <= cap) {
        // lists sorted on chip: registers (<= 32) or shared memory; one streaming pass in, one out
        static_assert(kWarpSortCap == 512, "k_emit_sorted splits its queues at 512");
        DevBuf<vid_t> big(n);
        DevBuf<int> nbig(2);
        nbig.zero();
        k_emit_sorted<<<grid_for(n * 32, 256), 256, 0, r.stream>>>(g.off.p, rnbr.p, n, rank_dev, doff.p, dnbr.p, big.p,
                                                                  nbig.p);
        launched();
        int h_nb[2];
        nbig.download(h_nb, 2);
        if (h_nb[0]) {
            const int grid = (int)std::min<int64_t>(ceil_div(h_nb[0], 8), (int64_t)r.sm_count * 12);
            k_sort_mid<<<grid, 256, 0, r.stream>>>(big.p, h_nb[0], doff.p, dnbr.p); launched();
        }
        if (h_nb[1]) {
            int P = 64;
            while (P < maxd) P <<= 1;
            const int grid = (int)std::min<int64_t>(h_nb[1], (int64_t)r.sm_count * 8);
            k_sort_big<<<grid, 256, (size_t)P * sizeof(vid_t), r.stream>>>(big.p + (n - h_nb[1]), h_nb[1], doff.p, dnbr.p);
            launched();
        }
    } else {
        // general fallback for lists longer than the on-chip sorter: one global radix sort of (rank[u], rank[v]) keys
        DevBuf<uint64_t> keys(m), alt(m);
        k_emit_out<<<grid_for(n * 32, 256), 256, 0, r.stream>>>(g.off.p, g.nbr.p, n, rank_dev, doff.p, keys.p);
        launched();
        uint64_t *sorted = radix_sort_keys(keys.p, alt.p, m, 0, 32 + bits_for((uint64_t)(n - 1)));
        k_low32<<<grid_for(m, 256), 256, 0, r.stream>>>(sorted, m, dnbr.p); launched();
    }
}

Dag *build_degree_dag(const Graph &g) {
    GMSB_REQUIRE(!g.directed, "degree orientation needs an undirected graph");
    auto *d = new Dag();
    try {
        d->n = g.n;
        degree_order(g, d->order, d->rank);
        orient_by_rank(g, d->rank.p, d->off, d->nbr, &d->m, &d->max_dplus);
    } catch (...) { delete d; throw; }
    return d;
}

// InduceDirectedGraph: the result is itself a (directed) Graph handle whose ids are ranks.
Graph *induce_directed(const Graph &g, const vid_t *ranking_host) {
    GMSB_REQUIRE(!g.directed, "Graph must be undirected");          // apply_order.h:14-16
    GMSB_REQUIRE(ranking_host != nullptr || g.n == 0, "orient: null ranking");
    Runtime &r = rt();
    int64_t n = g.n;
    DevBuf<vid_t> rank(n);
    rank.upload(ranking_host, n);
    if (n) {
        DevBuf<int> seen(n + 1);
        seen.zero();
        k_check_perm<<<grid_for(n, 256), 256, 0, r.stream>>>(rank.p, n, seen.p, seen.p + n); launched();
        GMSB_REQUIRE(seen.get(n) == 0, "orient: ranking is not a permutation of 0..n-1");
    }
    auto *out = new Graph();
    try {
        out->directed = true;
        int64_t m = 0;
        orient_by_rank(g, rank.p, out->off, out->nbr, &m, nullptr);
        out->slots = m;
        // n = max id that occurs in the induced edge list + 1 (builder.h:285); an empty list gives n = 1
        DevBuf<int> mx(1);
        mx.zero();
        if (n) { k_max_endpoint<<<grid_for(n, 256), 256, 0, r.stream>>>(out->off.p, out->nbr.p, n, mx.p); launched(); }
        out->n = (int64_t)mx.get(0) + 1;
        GMSB_CUDA(cudaStreamSynchronize(r.stream));
    } catch (...) { delete out; throw; }
    return out;
}

Dag::~Dag() { delete_plan(plan); }
Graph::~Graph() { delete dag; }

}  // namespace gmsb
