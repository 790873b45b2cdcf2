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
// d+(u) goes to dold[u] (indexed by the ORIGINAL id: every pass up to the final row move works on vertex ranges of the
// input, so that ranges can be processed as they arrive from the host, or by different devices).  Vertices with a big
// list are queued (bigq) for k_relabel_big.
// (The pass runs over the vertex range [u_begin, n): the upload pipeline calls it range by range; ids are bounds-checked
// here because in that pipeline the validation of a range is only enqueued, not yet read back, when this runs.)
__global__ void __launch_bounds__(256)
k_relabel_count(const eid_t *__restrict__ off, const vid_t *__restrict__ nbr, int64_t u_begin, int64_t n, int64_t n_all,
                const vid_t *__restrict__ rank, vid_t *__restrict__ rnbr, int32_t *__restrict__ dold,
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
            else dold[u] = c;
        }
    }
}
// big lists: blockIdx.x = queue entry, blockIdx.y = one of gridDim.y interleaved parts of the list (dold starts at 0)
__global__ void __launch_bounds__(256)
k_relabel_big(const vid_t *__restrict__ bigq, const eid_t *__restrict__ off, const vid_t *__restrict__ nbr,
              const vid_t *__restrict__ rank, vid_t *__restrict__ rnbr, int32_t *__restrict__ dold, int64_t n_all) {
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
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&dold[u], c);
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

// (2) emit: the chunk's higher-ranked neighbours are compacted with ballots into the FRONT OF THE ROW'S OWN SLOT RANGE of
// a second slots-sized array (row u of the symmetric CSR starts at off[u]; its d+(u) survivors land at
// trow[off[u] .. off[u] + d+(u))), so the pass needs nothing from outside its vertex range — no offsets of the
// oriented graph, which exist only once every vertex has been counted.  Lists of <= kLaneSortMax survivors are staged in
// the warp's shared-memory slice, sorted there by their owner lane (insertion sort: 32 lists at once) and written in
// final order; longer ones are written unsorted and queued for the warp / CTA sorters as (start << 24 | length).
constexpr int kQLenBits = 24;
constexpr uint64_t kQLenMask = (1ull << kQLenBits) - 1ull;
__global__ void __launch_bounds__(256)
k_emit_rows(const eid_t *__restrict__ off, const vid_t *__restrict__ rnbr, int64_t u_begin, int64_t n,
            const vid_t *__restrict__ rank, const int32_t *__restrict__ dold, vid_t *__restrict__ trow,
            uint64_t *__restrict__ queue, int64_t qcap, int *__restrict__ nq) {
    __shared__ vid_t stage_s[8][32 * kLaneSortMax];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    vid_t *stage = stage_s[wib];
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u0 = u_begin + warp * 32; u0 < n; u0 += nwarps * 32) {
        const WarpChunk ck = load_chunk(off, u0, n, lane);
        const int64_t u = u0 + lane;
        const bool mine = u < n && !ck.big;
        const vid_t ru = u < n ? rank[u] : 0;
        const eid_t ob = ck.abs;                                     // the row's survivors start where the row starts
        const int c = mine ? dold[u] : 0;
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
            const eid_t oob = __shfl_sync(0xffffffffu, ob, owner);
            const eid_t s = oob + (i - orel);
            const vid_t ro = __shfl_sync(0xffffffffu, ru, owner);
            const vid_t rv = act ? rnbr[s] : -1;
            const bool kp = act && rv > ro;
            const unsigned kept = __ballot_sync(0xffffffffu, kp);
            const int olo = max(orel - i0, 0);                       // first lane of the owner's segment in this window
            const int before = __popc(kept & ((1u << lane) - 1u) & ~((1u << olo) - 1u));
            const int pos = __shfl_sync(0xffffffffu, w, owner) + before;
            const bool osmall = __shfl_sync(0xffffffffu, (int)small, owner) != 0;
            const int osp = __shfl_sync(0xffffffffu, sp, owner);
            if (kp) {
                if (osmall) stage[osp + pos] = rv; else trow[oob + pos] = rv;
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
            if (act) trow[oob + (i - osp)] = stage[i];
        }
        if (mine && !small) {
            // two queues filled from opposite ends of one array: [0, nq[0]) mid-size lists, (qcap-1-nq[1], qcap-1] long
            const uint64_t entry = ((uint64_t)ob << kQLenBits) | (uint64_t)c;
            if (c <= 512) queue[atomicAdd(&nq[0], 1)] = entry;
            else queue[qcap - 1 - atomicAdd(&nq[1], 1)] = entry;
        }
    }
}
// big lists: survivors appended through one cursor per list (unsorted; the sorters fix the order)
__global__ void __launch_bounds__(256)
k_emit_big(const vid_t *__restrict__ bigq, const eid_t *__restrict__ off, const vid_t *__restrict__ rnbr,
           const vid_t *__restrict__ rank, const int32_t *__restrict__ dold, vid_t *__restrict__ trow,
           int *__restrict__ cursor, uint64_t *__restrict__ queue, int64_t qcap, int *__restrict__ nq) {
    const vid_t u = bigq[blockIdx.x];
    const eid_t b = off[u], e = off[u + 1];
    const vid_t ru = rank[u];
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
        if (kp) trow[b + base + __popc(kept & ((1u << lane) - 1u))] = rv;
    }
    if (blockIdx.y == 0 && threadIdx.x == 0) {
        const int c = dold[u];
        if (c > 1) {
            const uint64_t entry = ((uint64_t)b << kQLenBits) | (uint64_t)c;
            if (c <= 512) queue[atomicAdd(&nq[0], 1)] = entry;
            else queue[qcap - 1 - atomicAdd(&nq[1], 1)] = entry;
        }
    }
}

// (3a) warp sorter for queued lists of up to kWarpSortCap elements: one warp per list; up to 32 elements in registers
// (shuffle network), longer ones as a bitonic network in the warp's own shared-memory slice (no block barriers), eight
// lists per CTA in flight.
constexpr int kWarpSortCap = 512;
__global__ void __launch_bounds__(256)
k_sort_mid(const uint64_t *__restrict__ queue, int nlists, vid_t *__restrict__ rows) {
    __shared__ vid_t bufs[8][kWarpSortCap];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    vid_t *buf = bufs[wib];
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < nlists; t += nwarps) {
        const uint64_t entry = queue[t];
        const eid_t ob = (eid_t)(entry >> kQLenBits);
        const int c = (int)(entry & kQLenMask);
        if (c > kWarpSortCap) continue;                    // left to the CTA sorter
        if (c <= 32) {
            const vid_t v = warp_bitonic_sort(lane < c ? rows[ob + lane] : 0x7fffffff, lane);
            if (lane < c) rows[ob + lane] = v;
            continue;
        }
        int P = 64;
        while (P < c) P <<= 1;
        __syncwarp();
        for (int i = lane; i < P; i += 32) buf[i] = i < c ? rows[ob + i] : 0x7fffffff;
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
        for (int i = lane; i < c; i += 32) rows[ob + i] = buf[i];
    }
}

// (3b) CTA sorter for the few longer lists: bitonic network in shared memory (kSortCap elements).
constexpr int kSortCap = 8192;
__global__ void __launch_bounds__(256)
k_sort_big(const uint64_t *__restrict__ queue, int nlists, vid_t *__restrict__ rows) {
    extern __shared__ vid_t buf[];
    for (int t = blockIdx.x; t < nlists; t += gridDim.x) {
        const uint64_t entry = queue[t];
        const eid_t ob = (eid_t)(entry >> kQLenBits);
        const int c = (int)(entry & kQLenMask);
        if (c <= kWarpSortCap || c > kSortCap) continue;   // done by the warp sorter / left to the radix-sort fallback
        int P = 64;
        while (P < c) P <<= 1;
        __syncthreads();
        for (int i = threadIdx.x; i < P; i += blockDim.x) buf[i] = i < c ? rows[ob + i] : 0x7fffffff;
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
        for (int i = threadIdx.x; i < c; i += blockDim.x) rows[ob + i] = buf[i];
    }
}

// (4) offsets of the oriented graph: d+ scattered from original-id order to rank order (an inclusive scan follows)
__global__ void k_counts_by_rank(const int32_t *__restrict__ dold, const vid_t *__restrict__ rank, int64_t n,
                                 eid_t *__restrict__ cnt /* n+1, [0] = 0 */) {
    for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < n; u += (int64_t)gridDim.x * blockDim.x)
        cnt[rank[u] + 1] = (eid_t)dold[u];
    if (blockIdx.x == 0 && threadIdx.x == 0) cnt[0] = 0;
}

// (5) row move: the len[u] elements at src[src_start[u] ..) go to dst[dst_start ..) for every vertex u of the range,
// where dst_start = dst_off[rank[u]] (BY_RANK: rows into their place in the oriented CSR) or dst_off[u - u_begin] (rows
// packed in original-id order: the piece a device hands to the others).  Warp chunks as above: 32 rows per warp, their
// elements streamed 32 at a time.
template <bool BY_RANK>
__global__ void __launch_bounds__(256)
k_move_rows(const eid_t *__restrict__ src_start, const int32_t *__restrict__ len, const vid_t *__restrict__ src,
            int64_t u_begin, int64_t u_end, const eid_t *__restrict__ dst_off, const vid_t *__restrict__ rank,
            vid_t *__restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u0 = u_begin + warp * 32; u0 < u_end; u0 += nwarps * 32) {
        const int64_t u = u0 + lane;
        const bool valid = u < u_end;
        const int c = valid ? len[u] : 0;
        const eid_t ss = valid ? src_start[u] : 0;
        const eid_t dd = valid ? dst_off[BY_RANK ? (int64_t)rank[u] : u - u_begin] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += x;
        }
        const int rel = incl - c;
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        for (int i0 = 0; i0 < total; i0 += 32) {
            const int i = i0 + lane;
            const bool act = i < total;
            const int owner = chunk_owner(rel, act ? i : total - 1);
            const int j = i - __shfl_sync(0xffffffffu, rel, owner);
            const eid_t so = __shfl_sync(0xffffffffu, ss, owner), dofs = __shfl_sync(0xffffffffu, dd, owner);
            if (act) dst[dofs + j] = src[so + j];
        }
    }
}

// where vertex u's row sits in the gathered pieces of a sharded build: piece r holds the rows of the original ids
// [cut[r], cut[r+1]) packed in that order and starts at r * stride; scan = exclusive scan of d+ in original-id order
constexpr int kMaxParts = 64;
struct PartCuts { int64_t cut[kMaxParts + 1]; int parts; };
__global__ void k_piece_starts(const eid_t *__restrict__ scan, int64_t n, PartCuts pc, int64_t stride,
                               eid_t *__restrict__ src_start) {
    for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < n; u += (int64_t)gridDim.x * blockDim.x) {
        int r = 0;
        while (r + 1 < pc.parts && u >= pc.cut[r + 1]) ++r;
        src_start[u] = scan[u] - scan[pc.cut[r]] + (eid_t)r * stride;
    }
}

__global__ void k_max_i32(const int32_t *__restrict__ a, int64_t n, int *out) {
    int mx = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        mx = max(mx, a[i]);
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(out, mx);
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

// ---- host side of the row-local orientation ---------------------------------------------------------------------------
// The passes are driven in three ways: over the whole graph at once (build_degree_dag, induce_directed), vertex range by
// vertex range while the host arrays are still uploading (graph_build.cu), and over one vertex range per device with the
// finished rows exchanged between the devices (shard_*).
void orient_rows_begin(const Graph &g, OrientRows &w) {
    w.rnbr.alloc(g.slots);
    w.trow.alloc(g.slots);
    w.dold.alloc(g.n + 1);                      // one spare entry: scans over a range read one element past it
    w.dold.zero();
    w.bigq.alloc(g.n);
    w.queue.alloc(g.n);                         // at most one entry per vertex
    w.counters.alloc(3);                        // [0] big lists queued  [1] mid sort queue  [2] long sort queue
    w.counters.zero();
    w.big_done = w.sorted_mid = w.sorted_long = 0;
}

// passes (1) + (2) over the vertex range [u_begin, u_end): needs the offsets and the neighbour slots of the range
void orient_rows_range(const Graph &g, const vid_t *rank_dev, OrientRows &w, int64_t u_begin, int64_t u_end) {
    Runtime &r = rt();
    if (u_end <= u_begin) return;
    const int grid = grid_for(u_end - u_begin, 256);
    k_relabel_count<<<grid, 256, 0, r.stream>>>(g.off.p, g.nbr.p, u_begin, u_end, g.n, rank_dev, w.rnbr.p, w.dold.p,
                                                w.bigq.p, w.counters.p);
    launched();
    k_emit_rows<<<grid, 256, 0, r.stream>>>(g.off.p, w.rnbr.p, u_begin, u_end, rank_dev, w.dold.p, w.trow.p, w.queue.p,
                                            (int64_t)w.queue.n, w.counters.p + 1);
    launched();
}

// lists of more than kBigList slots queued since the last call: many CTAs per list
void orient_rows_big(const Graph &g, const vid_t *rank_dev, OrientRows &w) {
    Runtime &r = rt();
    if (g.n == 0) return;
    const int queued = w.counters.get(0);
    const int fresh = queued - w.big_done;
    if (fresh <= 0) return;
    const vid_t *q = w.bigq.p + w.big_done;
    k_relabel_big<<<dim3((unsigned)fresh, kBigParts), 256, 0, r.stream>>>(q, g.off.p, g.nbr.p, rank_dev, w.rnbr.p,
                                                                         w.dold.p, g.n);
    launched();
    DevBuf<int> cursor(fresh);
    cursor.zero();
    k_emit_big<<<dim3((unsigned)fresh, kBigParts), 256, 0, r.stream>>>(q, g.off.p, w.rnbr.p, rank_dev, w.dold.p, w.trow.p,
                                                                      cursor.p, w.queue.p, (int64_t)w.queue.n,
                                                                      w.counters.p + 1);
    launched();
    GMSB_CUDA(cudaStreamSynchronize(r.stream));     // cursor is released at the end of this scope
    w.big_done = queued;
}

// sorts the lists queued since the last call (reads the queue counters back: a host synchronisation)
void orient_rows_sort(OrientRows &w) {
    Runtime &r = rt();
    if (w.queue.n == 0) return;
    static_assert(kWarpSortCap == 512, "the emit kernels split their queues at 512");
    int h[3];
    w.counters.download(h, 3);
    const int mid = h[1] - w.sorted_mid, lng = h[2] - w.sorted_long;
    if (mid > 0) {
        const int grid = (int)std::min<int64_t>(ceil_div(mid, 8), (int64_t)r.sm_count * 12);
        k_sort_mid<<<grid, 256, 0, r.stream>>>(w.queue.p + w.sorted_mid, mid, w.trow.p); launched();
    }
    if (lng > 0) {              // the long queue grows downwards from the end of the array
        const int grid = (int)std::min<int64_t>(lng, (int64_t)r.sm_count * 6);
        k_sort_big<<<grid, 256, (size_t)kSortCap * sizeof(vid_t), r.stream>>>(w.queue.p + (w.queue.n - (size_t)h[2]), lng,
                                                                             w.trow.p);
        launched();
    }
    w.sorted_mid = h[1];
    w.sorted_long = h[2];
}

// offsets of the oriented graph (rank order) from d+ in original-id order
static void offsets_by_rank(int64_t n, const int32_t *dold, const vid_t *rank_dev, DevBuf<eid_t> &doff, int64_t *m_out,
                            int *max_dplus, DevBuf<int32_t> *dplus) {
    Runtime &r = rt();
    doff.alloc(n + 1);
    int maxd = 0;
    int64_t m = 0;
    if (n == 0) doff.zero();
    else {
        k_counts_by_rank<<<grid_for(n, 256), 256, 0, r.stream>>>(dold, rank_dev, n, doff.p); launched();
        inclusive_sum_inplace(doff.p, n + 1);
        m = doff.get(n);
        DevBuf<int> mx(1);
        mx.zero();
        if (dplus) dplus->alloc(n);
        k_max_deg<<<grid_for(n, 256), 256, 0, r.stream>>>(doff.p, n, mx.p, dplus ? dplus->p : nullptr); launched();
        maxd = mx.get(0);
    }
    *m_out = m;
    if (max_dplus) *max_dplus = maxd;
}

static int on_chip_sort_cap() {
    int cap = kSortCap;
    if (const char *env = getenv("GMSB_ORIENT_SORT_CAP")) cap = std::min(kSortCap, std::max(0, atoi(env)));   // tests
    return cap;
}

// every range has been through orient_rows_range: big lists, offsets, remaining sorts, rows into rank order
void orient_rows_finish(const Graph &g, const vid_t *rank_dev, OrientRows &w, DevBuf<eid_t> &doff, DevBuf<vid_t> &dnbr,
                        int64_t *m_out, int *max_dplus, DevBuf<int32_t> *dplus, PhaseTrace &tr) {
    Runtime &r = rt();
    const int64_t n = g.n;
    orient_rows_big(g, rank_dev, w);
    int maxd = 0;
    offsets_by_rank(n, w.dold.p, rank_dev, doff, m_out, &maxd, dplus);
    if (max_dplus) *max_dplus = maxd;
    const int64_t m = *m_out;
    tr.mark("orient: relabel + count + emit");
    dnbr.alloc(m);
    if (m == 0) return;
    if (maxd <= on_chip_sort_cap()) {
        // lists sorted on chip: registers (<= 32) or shared memory; then one streaming pass moves the rows into place
        orient_rows_sort(w);
        tr.mark("orient: list sorts");
        k_move_rows<true><<<grid_for(n, 256), 256, 0, r.stream>>>(g.off.p, w.dold.p, w.trow.p, 0, n, doff.p, rank_dev,
                                                                 dnbr.p);
        launched();
        tr.mark("orient: rows into rank order");
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
    PhaseTrace tr("GMSB_TC_TRACE");
    OrientRows w;
    orient_rows_begin(g, w);
    orient_rows_range(g, rank_dev, w, 0, g.n);
    orient_rows_finish(g, rank_dev, w, doff, dnbr, m_out, max_dplus, dplus, tr);
    GMSB_CUDA(cudaStreamSynchronize(rt().stream));      // the work buffers are released when this returns
}

void orient_pipeline_begin(const Graph &g, OrientPipeline &p) {
    GMSB_REQUIRE(!g.directed, "degree orientation needs an undirected graph");
    p.d = new Dag();
    p.d->n = g.n;
    degree_order(g, p.d->order, p.d->rank, &p.d->max_deg);
    orient_rows_begin(g, p.w);
}
void orient_pipeline_range(const Graph &g, OrientPipeline &p, int64_t u_begin, int64_t u_end) {
    orient_rows_range(g, p.d->rank.p, p.w, u_begin, u_end);
    orient_rows_sort(p.w);          // the range's lists are sorted while the next range is still on its way
}
void orient_pipeline_finish(const Graph &g, OrientPipeline &p) {
    PhaseTrace tr("GMSB_TC_TRACE");
    orient_rows_finish(g, p.d->rank.p, p.w, p.d->off, p.d->nbr, &p.d->m, &p.d->max_dplus, &p.d->dplus, tr);
    GMSB_CUDA(cudaStreamSynchronize(rt().stream));
    p.w = OrientRows();
}

// ---- sharded build: one vertex range per device --------------------------------------------------------------------------
// After orient_rows_range over its range [u0, u1) a device packs its finished rows in original-id order (the "piece"),
// the pieces and the d+ values are exchanged by the caller (one all-gather, one all-reduce), and every device moves all
// rows into rank order.
int64_t orient_piece_layout(const Graph &g, const vid_t *rank_dev, OrientRows &w, int64_t u0, int64_t u1,
                            DevBuf<eid_t> &piece_off) {
    Runtime &r = rt();
    orient_rows_big(g, rank_dev, w);
    orient_rows_sort(w);
    piece_off.alloc((size_t)(u1 - u0) + 1);
    if (u1 > u0) {
        DevBuf<int> mx(1);
        mx.zero();
        k_max_i32<<<grid_for(u1 - u0, 256), 256, 0, r.stream>>>(w.dold.p + u0, u1 - u0, mx.p); launched();
        GMSB_REQUIRE(mx.get(0) <= kSortCap, "sharded build: an oriented list is longer than the on-chip sorter holds");
    }
    exclusive_sum(w.dold.p + u0, piece_off.p, (u1 - u0) + 1);       // reads dold[u1]: the spare entry when u1 == n
    return piece_off.get((size_t)(u1 - u0));
}
void orient_piece_export(const Graph &g, OrientRows &w, int64_t u0, int64_t u1, const DevBuf<eid_t> &piece_off,
                         vid_t *piece_dst, int32_t *dplus_all) {
    Runtime &r = rt();
    if (u1 > u0) {
        k_move_rows<false><<<grid_for(u1 - u0, 256), 256, 0, r.stream>>>(g.off.p, w.dold.p, w.trow.p, u0, u1, piece_off.p,
                                                                        nullptr, piece_dst);
        launched();
        GMSB_CUDA(cudaMemcpyAsync(dplus_all + u0, w.dold.p + u0, sizeof(int32_t) * (size_t)(u1 - u0),
                                  cudaMemcpyDeviceToDevice, r.stream));
    }
    GMSB_CUDA(cudaStreamSynchronize(r.stream));
}
void dag_from_pieces(Dag &d, const int32_t *dplus_all, const vid_t *pieces, int64_t stride, const int64_t *cut, int parts) {
    Runtime &r = rt();
    const int64_t n = d.n;
    GMSB_REQUIRE(parts >= 1 && parts <= kMaxParts, "sharded build: too many parts");
    offsets_by_rank(n, dplus_all, d.rank.p, d.off, &d.m, &d.max_dplus, &d.dplus);
    d.nbr.alloc(d.m);
    if (d.m == 0) return;
    DevBuf<int32_t> len(n + 1);
    len.zero();
    GMSB_CUDA(cudaMemcpyAsync(len.p, dplus_all, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToDevice, r.stream));
    DevBuf<eid_t> scan(n + 1), src_start(n);
    exclusive_sum(len.p, scan.p, n + 1);
    PartCuts pc{};
    pc.parts = parts;
    for (int i = 0; i <= parts; ++i) pc.cut[i] = cut[i];
    k_piece_starts<<<grid_for(n, 256), 256, 0, r.stream>>>(scan.p, n, pc, stride, src_start.p); launched();
    k_move_rows<true><<<grid_for(n, 256), 256, 0, r.stream>>>(src_start.p, len.p, pieces, 0, n, d.off.p, d.rank.p, d.nbr.p);
    launched();
    GMSB_CUDA(cudaStreamSynchronize(r.stream));
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
