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


// keys[v] = degree, ids[v] = v, and the largest degree (the number of key bits the sort has to look at)
__global__ void k_degree_keys(const eid_t *__restrict__ off, int64_t n, uint32_t *__restrict__ keys,
                              vid_t *__restrict__ ids, unsigned long long *__restrict__ maxdeg) {
    unsigned long long mx = 0;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long dg = (unsigned long long)(off[v + 1] - off[v]);
        keys[v] = (uint32_t)dg;
        ids[v] = (vid_t)v;
        mx = mx > dg ? mx : dg;
    }
    for (int o = 16; o; o >>= 1) { const unsigned long long x = __shfl_xor_sync(0xffffffffu, mx, o); mx = mx > x ? mx : x; }
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(maxdeg, mx);
}

__global__ void k_rank_from_order(const vid_t *__restrict__ order, int64_t n, vid_t *__restrict__ rank) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        rank[order[i]] = (vid_t)i;
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

__global__ void k_max_deg(const eid_t *__restrict__ off, int64_t n, int *out, int32_t *__restrict__ deg_out) {
    int mx = 0;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        const int dg = (int)(off[v + 1] - off[v]);
        if (deg_out) deg_out[v] = dg;
        mx = max(mx, dg);
    }
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
// Both passes walk the symmetric CSR as a WARP CHUNK: a warp takes 32 consecutive vertices, lane t keeps the start of
// vertex u0+t's list in the chunk's own slot numbering, and the warp streams the chunk's slots 32 at a time — slot i
// finds its owner with five shuffles (binary search over the 32 starts), the owner lanes count / place their own
// segment of each ballot.  The passes are therefore coalesced streams whose dependent-load chain is per chunk, not
// per vertex: round 1 gave every vertex its own warp iteration (16.7 M of them at scale 24, most with a handful of
// neighbours, 5.3 + 6.1 ms).  Lists longer than kBigList slots are left out of the chunk numbering and handled by
// their own launch, many CTAs per list, so that a hub's 400 K slots are not one warp's serial loop.
constexpr int kBigList = 4096;
constexpr int kLaneSortMax = 16;      // lists of <= 16 survivors are sorted by their owner lane in shared memory

struct WarpChunk {
    eid_t abs;      // off[u] of this lane's vertex
    int rel;        // start of the lane's list in the chunk numbering
    int deg;        // its length there (0 for lanes past n and for lists > kBigList)
    int total;      // slots of the chunk (same on every lane)
    bool big;       // this lane's list is > kBigList
};

__device__ __forceinline__ WarpChunk load_chunk(const eid_t *__restrict__ off, int64_t u0, int64_t n, int lane) {
    WarpChunk c;
    const int64_t u = u0 + lane;
    const eid_t b = u < n ? off[u] : 0, e = u < n ? off[u + 1] : 0;
    c.abs = b;
    c.big = e - b > kBigList;
    c.deg = c.big ? 0 : (int)(e - b);
    int incl = c.deg;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int x = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += x;
    }
    c.rel = incl - c.deg;
    c.total = __shfl_sync(0xffffffffu, incl, 31);
    return c;
}
// owner of chunk slot i: the largest lane t with rel_t <= i (rel_0 = 0; among equal starts the last one is the
// non-empty list)
__device__ __forceinline__ int chunk_owner(int rel, int i) {
    int pos = 0;
#pragma unroll
    for (int step = 16; step; step >>= 1) {
        const int oc = __shfl_sync(0xffffffffu, rel, pos + step);
        if (oc <= i) pos += step;
    }
    return pos;
}
// bits [lo, hi) of a 32-slot window that belong to the segment [rel, rel + deg) when the window starts at i0
__device__ __forceinline__ unsigned window_mask(int rel, int deg, int i0) {
    const int lo = min(max(rel - i0, 0), 32), hi = min(max(rel + deg - i0, 0), 32);
    const unsigned below_hi = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
    const unsigned below_lo = lo >= 32 ? 0xffffffffu : ((1u << lo) - 1u);
    return below_hi & ~below_lo;
}

// (1) relabel + count: rnbr[s] = rank[nbr[s]] is stored so that the emit pass streams instead of gathering again;
// d+(u) goes to cnt[rank[u] + 1].  Vertices with a big list are queued (bigq) for k_relabel_big.
// (The pass runs over the vertex range [u_begin, n): the upload pipeline calls it range by range; ids are bounds-checked
// here because in that pipeline the validation of a range is only enqueued, not yet read back, when this runs.)
__global__ void __launch_bounds__(256)
k_relabel_count(const eid_t *__restrict__ off, const vid_t *__restrict__ nbr, int64_t u_begin, int64_t n, int64_t n_all,
                const vid_t *__restrict__ rank, vid_t *__restrict__ rnbr, eid_t *__restrict__ cnt,
                vid_t *__restrict__ bigq, int *__restrict__ nbigq) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u0 = u_begin + warp * 32; u0 < n; u0 += nwarps * 32) {
        const WarpChunk ck = load_chunk(off, u0, n, lane);
        const int64_t u = u0 + lane;
        const vid_t ru = u < n ? rank[u] : 0;
        int c = 0;
#pragma unroll 2
        for (int i0 = 0; i0 < ck.total; i0 += 32) {
            const int i = i0 + lane;
            const bool act = i < ck.total;
            const int owner = chunk_owner(ck.rel, act ? i : ck.total - 1);
            const eid_t s = __shfl_sync(0xffffffffu, ck.abs, owner) + (i - __shfl_sync(0xffffffffu, ck.rel, owner));
            const vid_t ro = __shfl_sync(0xffffffffu, ru, owner);
            vid_t rv = -1;
            if (act) {
                const vid_t v = nbr[s];
                rv = (uint64_t)v < (uint64_t)n_all ? rank[v] : -1;
                rnbr[s] = rv;
            }
            const unsigned kept = __ballot_sync(0xffffffffu, act && rv > ro);
            c += __popc(kept & window_mask(ck.rel, ck.deg, i0));
        }
        if (u < n) {
            if (ck.big) bigq[atomicAdd(nbigq, 1)] = (vid_t)u;
            else cnt[ru + 1] = c;
        }
    }
}
// big lists: blockIdx.x = queue entry, blockIdx.y = one of gridDim.y interleaved parts of the list
__global__ void __launch_bounds__(256)
k_relabel_big(const vid_t *__restrict__ bigq, const eid_t *__restrict__ off, const vid_t *__restrict__ nbr,
              const vid_t *__restrict__ rank, vid_t *__restrict__ rnbr, eid_t *__restrict__ cnt, int64_t n_all) {
    const vid_t u = bigq[blockIdx.x];
    const eid_t b = off[u], e = off[u + 1];
    const vid_t ru = rank[u];
    int c = 0;
    for (eid_t s = b + (eid_t)blockIdx.y * blockDim.x + threadIdx.x; s < e; s += (eid_t)gridDim.y * blockDim.x) {
        const vid_t v = nbr[s];
        const vid_t rv = (uint64_t)v < (uint64_t)n_all ? rank[v] : -1;
        rnbr[s] = rv;
        c += rv > ru;
    }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(reinterpret_cast<unsigned long long *>(&cnt[ru + 1]), (unsigned long long)c);
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

// (2) emit: the chunk's higher-ranked neighbours are compacted with ballots.  Lists of <= kLaneSortMax survivors are
// staged in the warp's shared-memory slice, sorted there by their owner lane (insertion sort: 32 lists at once) and
// written in final order; longer ones are written unsorted and queued for the warp / CTA sorters.
__global__ void __launch_bounds__(256)
k_emit_sorted(const eid_t *__restrict__ off, const vid_t *__restrict__ rnbr, int64_t n, const vid_t *__restrict__ rank,
              const eid_t *__restrict__ doff, vid_t *__restrict__ dnbr, vid_t *__restrict__ big, int *__restrict__ nbig) {
    __shared__ vid_t stage_s[8][32 * kLaneSortMax];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    vid_t *stage = stage_s[wib];
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u0 = warp * 32; u0 < n; u0 += nwarps * 32) {
        const WarpChunk ck = load_chunk(off, u0, n, lane);
        const int64_t u = u0 + lane;
        const bool mine = u < n && !ck.big;
        const vid_t ru = u < n ? rank[u] : 0;
        const eid_t ob = mine ? doff[ru] : 0;
        const int c = mine ? (int)(doff[ru + 1] - ob) : 0;
        const bool small = c <= kLaneSortMax;
        const int sc = small ? c : 0;
        int sincl = sc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, sincl, o);
            if (lane >= o) sincl += x;
        }
        const int sp = sincl - sc;                                   // this lane's slice of the staging area
        const int nstaged = __shfl_sync(0xffffffffu, sincl, 31);
        int w = 0;                                                   // survivors of this lane's list placed so far
        __syncwarp();
        for (int i0 = 0; i0 < ck.total; i0 += 32) {
            const int i = i0 + lane;
            const bool act = i < ck.total;
            const int owner = chunk_owner(ck.rel, act ? i : ck.total - 1);
            const int orel = __shfl_sync(0xffffffffu, ck.rel, owner);
            const eid_t s = __shfl_sync(0xffffffffu, ck.abs, owner) + (i - orel);
            const vid_t ro = __shfl_sync(0xffffffffu, ru, owner);
            const vid_t rv = act ? rnbr[s] : -1;
            const bool kp = act && rv > ro;
            const unsigned kept = __ballot_sync(0xffffffffu, kp);
            const int olo = max(orel - i0, 0);                       // first lane of the owner's segment in this window
            const int before = __popc(kept & ((1u << lane) - 1u) & ~((1u << olo) - 1u));
            const int pos = __shfl_sync(0xffffffffu, w, owner) + before;
            const bool osmall = __shfl_sync(0xffffffffu, (int)small, owner) != 0;
            const int osp = __shfl_sync(0xffffffffu, sp, owner);
            const eid_t oob = __shfl_sync(0xffffffffu, ob, owner);
            if (kp) {
                if (osmall) stage[osp + pos] = rv; else dnbr[oob + pos] = rv;
            }
            w += __popc(kept & window_mask(ck.rel, ck.deg, i0));
        }
        __syncwarp();
        if (small && c > 1) {                                        // insertion sort of the lane's own list
            vid_t *a = stage + sp;
            for (int x = 1; x < c; ++x) {
                const vid_t key = a[x];
                int y = x - 1;
                while (y >= 0 && a[y] > key) { a[y + 1] = a[y]; --y; }
                a[y + 1] = key;
            }
        }
        __syncwarp();
        for (int i = lane; i < ((nstaged + 31) & ~31); i += 32) {    // staged lists out, element-parallel
            const bool act = i < nstaged;
            const int owner = chunk_owner(sp, act ? i : nstaged - 1);
            const eid_t oob = __shfl_sync(0xffffffffu, ob, owner);
            const int osp = __shfl_sync(0xffffffffu, sp, owner);
            if (act) dnbr[oob + (i - osp)] = stage[i];
        }
        if (mine && !small) {
            // two queues filled from opposite ends of one array: [0, nbig[0]) mid-size lists, (n-1-nbig[1], n-1] long
            if (c <= 512) big[atomicAdd(&nbig[0], 1)] = ru;
            else big[n - 1 - atomicAdd(&nbig[1], 1)] = ru;
        }
    }
}
// big lists: survivors appended through one cursor per list (unsorted; the sorters fix the order)
__global__ void __launch_bounds__(256)
k_emit_big(const vid_t *__restrict__ bigq, const eid_t *__restrict__ off, const vid_t *__restrict__ rnbr,
           const vid_t *__restrict__ rank, const eid_t *__restrict__ doff, vid_t *__restrict__ dnbr,
           int *__restrict__ cursor, int64_t n, vid_t *__restrict__ big, int *__restrict__ nbig) {
    const vid_t u = bigq[blockIdx.x];
    const eid_t b = off[u], e = off[u + 1];
    const vid_t ru = rank[u];
    const eid_t ob = doff[ru];
    const int lane = threadIdx.x & 31;
    const eid_t first = b + (eid_t)blockIdx.y * blockDim.x, stride = (eid_t)gridDim.y * blockDim.x;
    for (eid_t s0 = first; s0 < e; s0 += stride) {                  // s0 is uniform over the CTA
        const eid_t s = s0 + threadIdx.x;
        const vid_t rv = s < e ? rnbr[s] : -1;
        const bool kp = rv > ru;
        const unsigned kept = __ballot_sync(0xffffffffu, kp);
        int base = 0;
        if (lane == 0 && kept) base = atomicAdd(&cursor[blockIdx.x], __popc(kept));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (kp) dnbr[ob + base + __popc(kept & ((1u << lane) - 1u))] = rv;
    }
    if (blockIdx.y == 0 && threadIdx.x == 0) {
        const int c = (int)(doff[ru + 1] - ob);
        if (c > 1) {
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

// (degree asc, id asc): the ids start out ascending and the LSD radix sort is stable, so sorting (degree, id) pairs
// on the bits of the largest degree alone gives the order — 3 passes over 8 B per vertex at scale 24 instead of the
// 8 passes over a 64-bit (degree << 32 | id) key that round 1 did.
void degree_order(const Graph &g, DevBuf<vid_t> &order, DevBuf<vid_t> &rank, int64_t *max_deg) {
    Runtime &r = rt();
    int64_t n = g.n;
    order.alloc(n);
    rank.alloc(n);
    if (max_deg) *max_deg = 0;
    if (n == 0) return;
    GMSB_REQUIRE(g.slots < (int64_t(1) << 32), "degree_order: degrees above 2^32 are not supported");
    DevBuf<uint32_t> keys(n), keys_alt(n);
    DevBuf<vid_t> ids(n);
    DevBuf<unsigned long long> mx(1);
    mx.zero();
    k_degree_keys<<<grid_for(n, 256), 256, 0, r.stream>>>(g.off.p, n, keys.p, ids.p, mx.p); launched();
    const unsigned long long maxdeg = mx.get(0);
    if (max_deg) *max_deg = (int64_t)maxdeg;
    vid_t *sorted = radix_sort_pairs32(keys.p, keys_alt.p, ids.p, order.p, n, bits_for(maxdeg));
    if (sorted != order.p)
        GMSB_CUDA(cudaMemcpyAsync(order.p, sorted, sizeof(vid_t) * n, cudaMemcpyDeviceToDevice, r.stream));
    k_rank_from_order<<<grid_for(n, 256), 256, 0, r.stream>>>(order.p, n, rank.p); launched();
    GMSB_CUDA(cudaStreamSynchronize(r.stream));        // ids / keys are released when this returns
}

constexpr int kBigParts = 32;                  // CTAs per big list

// stage 1 of orient_by_rank over the vertex range [u_begin, u_end): needs the offsets and the slots of the range
static void relabel_stage(const Graph &g, const vid_t *rank_dev, vid_t *rnbr, eid_t *doff, vid_t *bigq, int *nbigq,
                          int64_t u_begin, int64_t u_end) {
    Runtime &r = rt();
    if (u_end <= u_begin) return;
    k_relabel_count<<<grid_for(u_end - u_begin, 256), 256, 0, r.stream>>>(g.off.p, g.nbr.p, u_begin, u_end, g.n, rank_dev,
                                                                        rnbr, doff, bigq, nbigq);
    launched();
}

// stage 2: big lists, offsets, emit + sort
static void finish_stage(const Graph &g, const vid_t *rank_dev, DevBuf<vid_t> &rnbr, DevBuf<vid_t> &bigq,
                         DevBuf<int> &nbigq, DevBuf<eid_t> &doff, DevBuf<vid_t> &dnbr, int64_t *m_out, int *max_dplus,
                         DevBuf<int32_t> *dplus, PhaseTrace &tr) {
    Runtime &r = rt();
    const int64_t n = g.n;
    int n_bigq = 0;
    if (n) {
        n_bigq = nbigq.get(0);
        if (n_bigq) {
            k_relabel_big<<<dim3((unsigned)n_bigq, kBigParts), 256, 0, r.stream>>>(bigq.p, g.off.p, g.nbr.p, rank_dev,
                                                                                 rnbr.p, doff.p, n);
            launched();
        }
        inclusive_sum_inplace(doff.p, n + 1);
    }
    tr.mark("orient: relabel + count");
    int64_t m = n ? doff.get(n) : 0;
    *m_out = m;
    dnbr.alloc(m);
    int maxd = 0;
    if (n) {
        DevBuf<int> mx(1);
        mx.zero();
        if (dplus) dplus->alloc(n);
        k_max_deg<<<grid_for(n, 256), 256, 0, r.stream>>>(doff.p, n, mx.p, dplus ? dplus->p : nullptr); launched();
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
        k_emit_sorted<<<grid_for(n, 256), 256, 0, r.stream>>>(g.off.p, rnbr.p, n, rank_dev, doff.p, dnbr.p, big.p,
                                                             nbig.p);
        launched();
        if (n_bigq) {
            DevBuf<int> cursor(n_bigq);
            cursor.zero();
            k_emit_big<<<dim3((unsigned)n_bigq, kBigParts), 256, 0, r.stream>>>(bigq.p, g.off.p, rnbr.p, rank_dev, doff.p,
                                                                              dnbr.p, cursor.p, n, big.p, nbig.p);
            launched();
            GMSB_CUDA(cudaStreamSynchronize(r.stream));     // cursor is released at the end of this scope
        }
        tr.mark("orient: emit");
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
        tr.mark("orient: list sorts");
    } else {
        // general fallback for lists longer than the on-chip sorter: one global radix sort of (rank[u], rank[v]) keys
        DevBuf<uint64_t> keys(m), alt(m);
        k_emit_out<<<grid_for(n * 32, 256), 256, 0, r.stream>>>(g.off.p, g.nbr.p, n, rank_dev, doff.p, keys.p);
        launched();
        uint64_t *sorted = radix_sort_keys(keys.p, alt.p, m, 0, 32 + bits_for((uint64_t)(n - 1)));
        k_low32<<<grid_for(m, 256), 256, 0, r.stream>>>(sorted, m, dnbr.p); launched();
    }
}

void orient_by_rank(const Graph &g, const vid_t *rank_dev, DevBuf<eid_t> &doff, DevBuf<vid_t> &dnbr, int64_t *m_out,
                    int *max_dplus, DevBuf<int32_t> *dplus) {
    const int64_t n = g.n;
    PhaseTrace tr("GMSB_TC_TRACE");
    doff.alloc(n + 1);
    doff.zero();
    DevBuf<vid_t> rnbr(g.slots), bigq(n);
    DevBuf<int> nbigq(1);
    nbigq.zero();
    relabel_stage(g, rank_dev, rnbr.p, doff.p, bigq.p, nbigq.p, 0, n);
    finish_stage(g, rank_dev, rnbr, bigq, nbigq, doff, dnbr, m_out, max_dplus, dplus, tr);
}

void orient_pipeline_begin(const Graph &g, OrientPipeline &p) {
    GMSB_REQUIRE(!g.directed, "degree orientation needs an undirected graph");
    p.d = new Dag();
    p.d->n = g.n;
    degree_order(g, p.d->order, p.d->rank, &p.d->max_deg);
    p.d->off.alloc(g.n + 1);
    p.d->off.zero();
    p.rnbr.alloc(g.slots);
    p.bigq.alloc(g.n);
    p.nbigq.alloc(1);
    p.nbigq.zero();
}
void orient_pipeline_range(const Graph &g, OrientPipeline &p, int64_t u_begin, int64_t u_end) {
    relabel_stage(g, p.d->rank.p, p.rnbr.p, p.d->off.p, p.bigq.p, p.nbigq.p, u_begin, u_end);
}
void orient_pipeline_finish(const Graph &g, OrientPipeline &p) {
    PhaseTrace tr("GMSB_TC_TRACE");
    finish_stage(g, p.d->rank.p, p.rnbr, p.bigq, p.nbigq, p.d->off, p.d->nbr, &p.d->m, &p.d->max_dplus, &p.d->dplus, tr);
    p.rnbr.release(); p.bigq.release(); p.nbigq.release();
}

Dag *build_degree_dag(const Graph &g) {
    GMSB_REQUIRE(!g.directed, "degree orientation needs an undirected graph");
    auto *d = new Dag();
    try {
        d->n = g.n;
        PhaseTrace tr("GMSB_TC_TRACE");
        degree_order(g, d->order, d->rank, &d->max_deg);
        tr.mark("orient: degree order");
        orient_by_rank(g, d->rank.p, d->off, d->nbr, &d->m, &d->max_dplus, &d->dplus);
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
Graph::~Graph() {
    if (replicas && release_replicas) release_replicas(replicas);
    delete dag;
}

}  // namespace gmsb
