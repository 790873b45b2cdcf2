// tc.cu — triangle counting on the degree-oriented DAG: schedule + the three intersection kernels.
//
// Replaces  GMS::TriangleCount::{Seq,Par}::count_total   gms/algorithms/set_based/triangle_count/parallel/total.h:8-24
// whose inner loop is  vec_set_intersect_count           gms/representations/sets/sorted_set_operations.h:45-71
//
// Formulation.  In rank space a triangle {u<v<w} is the wedge u->v, u->w closed by the edge v->w, so
//     TC = sum over oriented edges (u,v) of | suffix(N+(u), after v)  ∩  N+(v) |
// (the elements of N+(u) up to and including v can never be in N+(v), whose members are all > v).  Every edge is
// therefore a pair (suffix descriptor, v); descriptors are grouped by v so that one CTA can keep N+(v) on chip:
//
//   bitmap  — v is a hub: N+(v) becomes a bitmap over (v, last(N+(v))] in shared memory; the CTA streams the
//             suffixes of all in-neighbours with coalesced loads and does one bit probe per element.
//   gallop  — skewed light pair: each lane binary-searches one element of the shorter list in the longer one.
//   merge   — balanced light pair: warp-cooperative merge path, lists staged through shared memory, every lane
//             walks an equal share of the merge diagonal.
//
// The choice is made per edge on the device when the schedule ("plan") is built: hub-ness of v first, then the
// length ratio of the two lists.  Counts are integers, so every variant returns the same value.
#include "common.cuh"
#include "sort.cuh"
#include "orient.cuh"
#include "isect.cuh"
#include "tc_plan.cuh"

namespace gmsb {

void delete_plan(TcPlan *p) { delete p; }

namespace {

// ---- plan construction -----------------------------------------------------------------------------------------
// One warp per vertex u: a descriptor for every out-edge; edges that cannot close a triangle (empty suffix or
// sink v) get the sentinel key n<<2 and sort to the tail.
__global__ void k_emit_desc(const eid_t *__restrict__ off, const vid_t *__restrict__ nbr, int64_t n,
                            uint32_t *__restrict__ keys, uint64_t *__restrict__ vals,
                            unsigned long long *__restrict__ work /* n: wedges arriving at v */,
                            unsigned long long *__restrict__ vbytes /* n: algorithmic bytes arriving at v */,
                            unsigned long long *__restrict__ acc /* [0]=B_TC [1]=wedges [2]=kept [3]=kept bytes */) {
    int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long bytes = 0, wedges = 0, kept = 0, kbytes = 0;
    for (int64_t u = warp; u < n; u += nwarps) {
        eid_t b = off[u], e = off[u + 1];
        for (eid_t s = b + lane; s < e; s += 32) {
            vid_t v = nbr[s];
            eid_t dv = off[v + 1] - off[v];
            eid_t len = e - s - 1;
            unsigned long long eb = 4ull * (unsigned long long)((e - b) + dv);
            bytes += eb;
            bool kp = len > 0 && dv > 0;
            uint32_t cls = len <= kShortLen ? 0u : (len <= kMidLen ? 1u : 2u);
            keys[s] = kp ? (((uint32_t)v << kClassBits) | cls) : ((uint32_t)n << kClassBits);
            vals[s] = ((uint64_t)(s + 1) << kLenBits) | (uint64_t)len;
            if (kp) {
                atomicAdd(&work[v], (unsigned long long)len);
                atomicAdd(&vbytes[v], eb);
                wedges += len; kept++; kbytes += eb;
            }
        }
    }
    for (int o = 16; o; o >>= 1) {
        bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
        wedges += __shfl_xor_sync(0xffffffffu, wedges, o);
        kept += __shfl_xor_sync(0xffffffffu, kept, o);
        kbytes += __shfl_xor_sync(0xffffffffu, kbytes, o);
    }
    if (lane == 0 && (bytes | kept)) {
        atomicAdd(&acc[0], bytes); atomicAdd(&acc[1], wedges); atomicAdd(&acc[2], kept); atomicAdd(&acc[3], kbytes);
    }
}

__device__ __forceinline__ int64_t first_key_ge(const uint32_t *__restrict__ keys, int64_t lo, int64_t hi, uint32_t x) {
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// inoff[v] = first descriptor of vertex v (keys sorted ascending, length cnt); inoff[n] = cnt
__global__ void k_vertex_bounds(const uint32_t *__restrict__ keys, int64_t cnt, int64_t n, int64_t *__restrict__ inoff) {
    for (int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; x <= n; x += (int64_t)gridDim.x * blockDim.x)
        inoff[x] = first_key_ge(keys, 0, cnt, (uint32_t)x << kClassBits);
}

struct PlanParams {
    int variant;
    int hub_bits;
    long long hub_min_work;
    long long item_cost;
};

// Per vertex: hub or not, and into how many CTA items its descriptor group is cut.
__global__ void k_classify(const eid_t *__restrict__ off, const vid_t *__restrict__ nbr, int64_t n,
                           const int64_t *__restrict__ inoff, const unsigned long long *__restrict__ work,
                           const unsigned long long *__restrict__ vbytes,
                           unsigned long long *__restrict__ cls /* [0]=bitmap bytes [1]=bitmap wedges */,
                           PlanParams pp, int64_t *__restrict__ nitems /* n+1, exclusive-scanned later */,
                           int *__restrict__ max_span_words) {
    int mx = 0;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        int64_t cnt = inoff[v + 1] - inoff[v];
        int64_t k = 0;
        if (cnt > 0 && (pp.variant == GMSB_TC_AUTO || pp.variant == GMSB_TC_BITMAP)) {
            eid_t ob = off[v], oe = off[v + 1];
            int64_t span = (int64_t)nbr[oe - 1] - v;            // bit x = w - v - 1, x in [0, span)
            bool hub = span <= pp.hub_bits &&
                       (pp.variant == GMSB_TC_BITMAP || (long long)work[v] >= pp.hub_min_work);
            if (hub) {
                int64_t words = (span + 31) >> 5;
                long long setup = 16ll * (oe - ob);
                long long target = pp.item_cost > setup ? pp.item_cost : setup;
                long long cost = (long long)work[v] + 4ll * cnt;
                k = (cost + target - 1) / target;
                if (k < 1) k = 1;
                if (k > cnt) k = cnt;
                mx = max(mx, (int)words);
                atomicAdd(&cls[0], vbytes[v]);
                atomicAdd(&cls[1], work[v]);
            }
        }
        nitems[v] = k;
    }
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(max_span_words, mx);
}

__global__ void k_fill_items(int64_t n, const uint32_t *__restrict__ keys, const int64_t *__restrict__ inoff,
                             const int64_t *__restrict__ nitems, const int64_t *__restrict__ item_base,
                             Item *__restrict__ items) {
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        int64_t k = nitems[v];
        if (!k) continue;
        int64_t b = inoff[v], e = inoff[v + 1], cnt = e - b;
        int64_t c0 = first_key_ge(keys, b, e, ((uint32_t)v << kClassBits) | 1u);     // end of class 0
        int64_t c1 = first_key_ge(keys, c0, e, ((uint32_t)v << kClassBits) | 2u);    // end of class 1
        int64_t chunk = (cnt + k - 1) / k;
        int64_t w = item_base[v];
        for (int64_t j = 0; j < k; ++j) {
            int64_t s = b + j * chunk;
            int64_t c = e - s < chunk ? e - s : chunk;
            if (c < 0) c = 0;
            Item it;
            it.v = (int32_t)v; it.begin = s; it.count = (int32_t)c;
            int64_t a0 = c0 - s, a1 = c1 - s;
            it.n0 = (int32_t)(a0 < 0 ? 0 : (a0 > c ? c : a0));
            it.n1 = (int32_t)(a1 < 0 ? 0 : (a1 > c ? c : a1));
            items[w + j] = it;
        }
    }
}

// Order key of an item: L2 tile of its first suffix (tile-major order keeps the lists that concurrently running
// CTAs stream inside the L2), then heaviest vertex first.
__global__ void k_item_keys(const Item *__restrict__ items, int64_t cnt, const uint64_t *__restrict__ desc, int64_t n,
                            int tile_shift, const eid_t *__restrict__ off, const vid_t *__restrict__ nbr,
                            uint64_t *__restrict__ keys, unsigned long long *__restrict__ n_small,
                            int *__restrict__ small_words) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cnt; i += (int64_t)gridDim.x * blockDim.x) {
        const Item it = items[i];
        // a slice that mixes classes has no single u-range: use its longest-suffix part (the bulk of its work)
        const int64_t d = it.begin + (it.n1 < it.count ? it.n1 : (it.n0 < it.count ? it.n0 : 0));
        const uint64_t tile = tile_shift > 0 ? (desc[d] >> kLenBits) >> tile_shift : 0;
        const int words = (int)(((int64_t)nbr[off[it.v + 1] - 1] - it.v + 31) >> 5);
        const bool wide = (words + 1) * 4 > kSmallWindowBytes;
        if (!wide) { atomicAdd(n_small, 1ull); atomicMax(small_words, words); }
        keys[i] = ((uint64_t)wide << 63) | (tile << 32) | (uint64_t)(uint32_t)(n - 1 - it.v);
    }
}

// Light descriptors: flag for merge / gallop by the length ratio of suffix and N+(v).
__global__ void k_flag_light(const uint32_t *__restrict__ keys, const uint64_t *__restrict__ vals, int64_t cnt,
                             const eid_t *__restrict__ off, const int64_t *__restrict__ nitems, int variant,
                             int ratio, uint8_t *__restrict__ fm, uint8_t *__restrict__ fg) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cnt; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t v = keys[i] >> kClassBits;
        bool light = nitems[v] == 0;
        bool gallop = false;
        if (light) {
            long long a = (long long)(vals[i] & kLenMask), b = off[v + 1] - off[v];
            long long lo = a < b ? a : b, hi = a < b ? b : a;
            gallop = variant == GMSB_TC_GALLOP || (variant != GMSB_TC_MERGE && hi >= (long long)ratio * lo);
        }
        fm[i] = light && !gallop;
        fg[i] = light && gallop;
    }
}

__global__ void k_compact_light(const uint32_t *__restrict__ keys, const uint64_t *__restrict__ vals, int64_t cnt,
                                const uint8_t *__restrict__ flag, const int64_t *__restrict__ pos,
                                uint64_t *__restrict__ odesc, vid_t *__restrict__ ov) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cnt; i += (int64_t)gridDim.x * blockDim.x)
        if (flag[i]) { odesc[pos[i]] = vals[i]; ov[pos[i]] = (vid_t)(keys[i] >> kClassBits); }
}

// ---- counting kernels ----------------------------------------------------------------------------------------------
// Hub bitmap kernel.  Persistent CTAs (one wave, grid = SMs x resident CTAs) pull items (v, slice of v's incoming
// suffix descriptors) from a global ticket, heaviest first.  The shared-memory bitmap is zeroed ONCE per CTA; each
// item sets the bits of N+(v), streams its suffixes, then clears exactly the words it set — so the per-item cost
// is O(d+(v)) and independent of how wide the window (v, last(N+(v))] is.  Inside an item the warps pull
// descriptors from shared-memory tickets, one per length class (suffix lengths vary by 1000x): a lane, an 8-lane
// group or the whole warp takes one descriptor, so short suffixes do not pay a warp's worth of control overhead.
// List elements are read with coalesced, 4-deep predicated loads and cost ONE branch-free probe each:
// out-of-window elements (and predicated-off slots, x = ~0) are clamped onto the always-zero guard word.
constexpr int kDescChunk = 4;

// G lanes cooperate on one descriptor; the warp takes 32/G descriptors per ticket (4 when G == 32).
template <int G, bool DEEP = false>
__device__ __forceinline__ uint32_t stream_class(const uint64_t *__restrict__ dptr, int lo, int hi, int *ticket,
                                                 const vid_t *__restrict__ nbr, const uint32_t *bm, uint32_t base,
                                                 uint32_t cap_words, int lane) {
    uint32_t hits = 0;
    if (lo >= hi) return 0;
    constexpr int BATCH = G == 32 ? kDescChunk : 32 / G;
    for (;;) {
        int d0 = 0;
        if (lane == 0) d0 = atomicAdd(ticket, BATCH);
        d0 = __shfl_sync(0xffffffffu, d0, 0) + lo;
        if (d0 >= hi) break;
        if constexpr (G == 32) {
            const int nd = min(BATCH, hi - d0);
            const uint64_t mine = lane < nd ? __ldcs(&dptr[d0 + lane]) : 0ull;
            for (int k = 0; k < nd; ++k) {
                const uint64_t ds = __shfl_sync(0xffffffffu, mine, k);
                const vid_t *__restrict__ p = nbr + (ds >> kLenBits);
                const int len = (int)(ds & kLenMask);
                if constexpr (DEEP) {
                    for (int j = lane; j < len; j += 256) {       // 8 loads in flight per lane
                        uint32_t x[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) x[q] = j + 32 * q < len ? (uint32_t)p[j + 32 * q] - base : ~0u;
#pragma unroll
                        for (int q = 0; q < 8; ++q) hits += probe(bm, x[q], cap_words);
                    }
                } else {
                    for (int j = lane; j < len; j += 128) {
                        const uint32_t x0 = (uint32_t)p[j] - base;
                        const uint32_t x1 = j + 32 < len ? (uint32_t)p[j + 32] - base : ~0u;
                        const uint32_t x2 = j + 64 < len ? (uint32_t)p[j + 64] - base : ~0u;
                        const uint32_t x3 = j + 96 < len ? (uint32_t)p[j + 96] - base : ~0u;
                        hits += probe(bm, x0, cap_words) + probe(bm, x1, cap_words) + probe(bm, x2, cap_words) +
                                probe(bm, x3, cap_words);
                    }
                }
            }
        } else {
            const int idx = d0 + lane / G, sub = lane % G;
            const uint64_t ds = idx < hi ? __ldcs(&dptr[idx]) : 0ull;
            const vid_t *__restrict__ p = nbr + (ds >> kLenBits);
            const int len = (int)(ds & kLenMask);
            if constexpr (G == 1) {
#pragma unroll
                for (int j = 0; j < kShortLen; ++j)
                    hits += probe(bm, j < len ? (uint32_t)p[j] - base : ~0u, cap_words);
            } else {
                for (int j = sub; j < len; j += 4 * G) {
                    const uint32_t x0 = (uint32_t)p[j] - base;
                    const uint32_t x1 = j + G < len ? (uint32_t)p[j + G] - base : ~0u;
                    const uint32_t x2 = j + 2 * G < len ? (uint32_t)p[j + 2 * G] - base : ~0u;
                    const uint32_t x3 = j + 3 * G < len ? (uint32_t)p[j + 3 * G] - base : ~0u;
                    hits += probe(bm, x0, cap_words) + probe(bm, x1, cap_words) + probe(bm, x2, cap_words) +
                            probe(bm, x3, cap_words);
                }
            }
        }
    }
    return hits;
}

template <int BLOCK, int MINB, bool DEEP = false>
__global__ void __launch_bounds__(BLOCK, MINB)
k_tc_bitmap(const Item *__restrict__ items, int64_t first, int64_t stride, int64_t count, uint32_t cap_words,
            const eid_t *__restrict__ off, const vid_t *__restrict__ nbr, const uint64_t *__restrict__ desc,
            unsigned long long *__restrict__ total, unsigned int *__restrict__ ticket) {
    extern __shared__ uint32_t bm[];                  // cap_words + 1 words
    __shared__ unsigned long long red[BLOCK / 32];
    __shared__ unsigned int s_item;
    __shared__ int s_next[3];
    const int tid = threadIdx.x, lane = tid & 31;
    for (uint32_t i = tid; i <= cap_words; i += BLOCK) bm[i] = 0u;
    unsigned long long hits64 = 0;
    for (;;) {
        if (tid == 0) { s_item = atomicAdd(ticket, 1u); s_next[0] = 0; s_next[1] = 0; s_next[2] = 0; }
        __syncthreads();                              // ticket visible; previous item's clears done
        const int64_t it = (int64_t)s_item;
        if (it >= count) break;
        const Item item = items[first + it * stride];                     // plan order: L2 tile, then heaviest
        const vid_t v = item.v;
        const eid_t ob = off[v], oe = off[v + 1];
        const uint32_t base = (uint32_t)v + 1u;
        for (eid_t j = ob + tid; j < oe; j += BLOCK) {
            const uint32_t x = (uint32_t)nbr[j] - base;
            atomicOr(&bm[x >> 5], 1u << (x & 31));
        }
        __syncthreads();                              // bitmap of N+(v) complete
        const uint64_t *__restrict__ dptr = desc + item.begin;
        uint32_t hits = stream_class<32, DEEP>(dptr, item.n1, item.count, &s_next[2], nbr, bm, base, cap_words, lane);
        hits += stream_class<8>(dptr, item.n0, item.n1, &s_next[1], nbr, bm, base, cap_words, lane);
        hits += stream_class<1>(dptr, 0, item.n0, &s_next[0], nbr, bm, base, cap_words, lane);
        hits64 += hits;
        __syncthreads();                              // every probe of this item done
        for (eid_t j = ob + tid; j < oe; j += BLOCK) bm[((uint32_t)nbr[j] - base) >> 5] = 0u;
    }
    unsigned long long s = block_sum(hits64, red);
    if (tid == 0 && s) atomicAdd(total, s);
}

// Galloping kernel: one warp per light edge (isect.cuh: warp_gallop_count).
__global__ void __launch_bounds__(256)
k_tc_gallop(const uint64_t *__restrict__ desc, const vid_t *__restrict__ vs, int64_t first, int64_t stride,
            int64_t count, const eid_t *__restrict__ off, const vid_t *__restrict__ nbr,
            unsigned long long *__restrict__ total) {
    __shared__ unsigned long long red[8];
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    uint32_t hits = 0;
    for (int64_t k = warp; k < count; k += nwarps) {
        const int64_t i = first + k * stride;
        const uint64_t ds = desc[i];
        const vid_t v = vs[i];
        const eid_t ob = off[v];
        hits += warp_gallop_count(nbr + (ds >> kLenBits), (int)(ds & kLenMask), nbr + ob, (int)(off[v + 1] - ob), lane);
    }
    unsigned long long s = block_sum(hits, red);
    if (threadIdx.x == 0 && s) atomicAdd(total, s);
}

// Merge-path kernel: one warp per light edge (isect.cuh: warp_merge_count), lists staged through shared memory.
constexpr int kMergeWarps = 8;

__global__ void __launch_bounds__(kMergeWarps * 32)
k_tc_merge(const uint64_t *__restrict__ desc, const vid_t *__restrict__ vs, int64_t first, int64_t stride,
           int64_t count, const eid_t *__restrict__ off, const vid_t *__restrict__ nbr,
           unsigned long long *__restrict__ total) {
    __shared__ vid_t stage[kMergeWarps][kMergeTile + 2];
    __shared__ unsigned long long red[kMergeWarps];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    uint32_t hits = 0;
    for (int64_t k = warp; k < count; k += nwarps) {
        const int64_t i = first + k * stride;
        const uint64_t ds = desc[i];
        const vid_t v = vs[i];
        const eid_t ob = off[v];
        hits += warp_merge_count(nbr + (ds >> kLenBits), (int)(ds & kLenMask), nbr + ob, (int)(off[v + 1] - ob), lane,
                                 stage[wib]);
    }
    unsigned long long s = block_sum(hits, red);
    if (threadIdx.x == 0 && s) atomicAdd(total, s);
}

// ---- host side ---------------------------------------------------------------------------------------------------
gmsb_tc_options normalise(const gmsb_tc_options *in) {
    gmsb_tc_options o{};
    if (in) o = *in;
    if (o.part_count <= 0) { o.part_count = 1; o.part_index = 0; }
    if (o.hub_bitmap_bits <= 0) o.hub_bitmap_bits = 512 * 1024;      // 64 KB of shared memory
    if (o.gallop_ratio <= 0) o.gallop_ratio = 8;
    if (o.hub_min_work <= 0) o.hub_min_work = 1024;
    return o;
}

bool same_plan(const gmsb_tc_options &a, const gmsb_tc_options &b) {
    return a.variant == b.variant && a.hub_bitmap_bits == b.hub_bitmap_bits && a.gallop_ratio == b.gallop_ratio &&
           a.hub_min_work == b.hub_min_work && a.reserved[0] == b.reserved[0] && a.reserved[1] == b.reserved[1];
}

TcPlan *build_plan(const Dag &d, const gmsb_tc_options &opt) {
    Runtime &r = rt();
    auto *p = new TcPlan();
    try {
        p->opt = opt;
        const int64_t n = d.n, m = d.m;
        GMSB_REQUIRE(d.max_dplus < (1 << kLenBits), "tc: out-degree too large for the descriptor format");
        GMSB_REQUIRE(m < (int64_t(1) << (64 - kLenBits)), "tc: too many edges for the descriptor format");
        GMSB_REQUIRE(n < (int64_t(1) << (32 - kClassBits)), "tc: too many vertices for the 32-bit schedule key");
        size_t smem_cap = r.smem_optin ? r.smem_optin : 48 * 1024;
        int hub_bits = opt.hub_bitmap_bits;
        if ((size_t)hub_bits / 8 > smem_cap - 1024) hub_bits = (int)((smem_cap - 1024) * 8);
        if (m == 0 || n == 0) return p;

        p->keys_a.alloc(m); p->keys_b.alloc(m); p->vals_a.alloc(m); p->vals_b.alloc(m);
        DevBuf<unsigned long long> work(n), vbytes(n), acc(4), cls(2);
        work.zero(); vbytes.zero(); acc.zero(); cls.zero();
        k_emit_desc<<<grid_for(n * 32, 256), 256, 0, r.stream>>>(d.off.p, d.nbr.p, n, p->keys_a.p, p->vals_a.p,
                                                                 work.p, vbytes.p, acc.p);
        launched();
        radix_sort_pairs(p->keys_a.p, p->keys_b.p, p->vals_a.p, p->vals_b.p, m, bits_for((uint64_t)n) + kClassBits,
                         &p->sorted_keys, &p->sorted_vals);
        unsigned long long h_acc[4];
        acc.download(h_acc, 4);
        p->algorithmic_bytes = h_acc[0];
        p->wedges = h_acc[1];
        p->n_desc = (int64_t)h_acc[2];
        p->bytes_kept = h_acc[3];
        const int64_t cnt = p->n_desc;
        if (cnt == 0) return p;

        DevBuf<int64_t> inoff(n + 1), nitems(n + 1), item_base(n + 1);
        k_vertex_bounds<<<grid_for(n + 1, 256), 256, 0, r.stream>>>(p->sorted_keys, cnt, n, inoff.p); launched();
        DevBuf<int> mxw(1);
        mxw.zero(); nitems.zero();
        PlanParams pp{opt.variant, hub_bits, (long long)opt.hub_min_work,
                      opt.reserved[0] > 0 ? (long long)opt.reserved[0] : 262144ll};
        k_classify<<<grid_for(n, 256), 256, 0, r.stream>>>(d.off.p, d.nbr.p, n, inoff.p, work.p, vbytes.p, cls.p, pp,
                                                          nitems.p, mxw.p);
        launched();
        exclusive_sum(nitems.p, item_base.p, n + 1);
        p->n_items = item_base.get(n);
        p->max_span_words = mxw.get(0);
        unsigned long long h_cls[2];
        cls.download(h_cls, 2);
        p->bytes_bitmap = h_cls[0];
        p->wedges_bitmap = h_cls[1];
        if (p->n_items) {
            p->items.alloc(p->n_items);
            k_fill_items<<<grid_for(n, 256), 256, 0, r.stream>>>(n, p->sorted_keys, inoff.p, nitems.p, item_base.p,
                                                                p->items.p);
            launched();
            // order: tile-major, heaviest first inside a tile
            DevBuf<uint64_t> ik(p->n_items), ik2(p->n_items);
            DevBuf<Item> items2(p->n_items);
            const int tile_shift = opt.reserved[1] > 0 ? opt.reserved[1] : (opt.reserved[1] < 0 ? 0 : 24);
            DevBuf<unsigned long long> nsm(1);
            DevBuf<int> smw(1);
            nsm.zero(); smw.zero();
            k_item_keys<<<grid_for(p->n_items, 256), 256, 0, r.stream>>>(p->items.p, p->n_items, p->sorted_vals, n,
                                                                        tile_shift, d.off.p, d.nbr.p, ik.p, nsm.p, smw.p);
            launched();
            p->n_items_small = (int64_t)nsm.get(0);
            p->small_span_words = smw.get(0);
            size_t bytes = 0;
            GMSB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, ik.p, ik2.p, p->items.p, items2.p, p->n_items, 0,
                                                      64, r.stream));
            DevBuf<uint8_t> tmp(bytes);
            GMSB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, ik.p, ik2.p, p->items.p, items2.p, p->n_items, 0,
                                                      64, r.stream));
            r.launches += 9;
            GMSB_CUDA(cudaStreamSynchronize(r.stream));
            p->items = std::move(items2);
        }
        // light edges -> two compacted lists
        DevBuf<uint8_t> fm(cnt), fg(cnt);
        DevBuf<int64_t> pm(cnt + 1), pg(cnt + 1);
        k_flag_light<<<grid_for(cnt, 256), 256, 0, r.stream>>>(p->sorted_keys, p->sorted_vals, cnt, d.off.p, nitems.p,
                                                              opt.variant, opt.gallop_ratio, fm.p, fg.p);
        launched();
        exclusive_sum(fm.p, pm.p, cnt);
        exclusive_sum(fg.p, pg.p, cnt);
        p->n_merge = pm.get(cnt - 1) + fm.get(cnt - 1);
        p->n_gallop = pg.get(cnt - 1) + fg.get(cnt - 1);
        p->n_bitmap_edges = cnt - p->n_merge - p->n_gallop;
        if (p->n_merge) {
            p->m_desc.alloc(p->n_merge); p->m_v.alloc(p->n_merge);
            k_compact_light<<<grid_for(cnt, 256), 256, 0, r.stream>>>(p->sorted_keys, p->sorted_vals, cnt, fm.p, pm.p,
                                                                     p->m_desc.p, p->m_v.p);
            launched();
        }
        if (p->n_gallop) {
            p->g_desc.alloc(p->n_gallop); p->g_v.alloc(p->n_gallop);
            k_compact_light<<<grid_for(cnt, 256), 256, 0, r.stream>>>(p->sorted_keys, p->sorted_vals, cnt, fg.p, pg.p,
                                                                     p->g_desc.p, p->g_v.p);
            launched();
        }
        // the key halves are no longer needed once the lists are built
        p->keys_a.release(); p->keys_b.release();
        if (p->sorted_vals == p->vals_a.p) p->vals_b.release(); else p->vals_a.release();
        GMSB_CUDA(cudaStreamSynchronize(r.stream));
    } catch (...) { delete p; throw; }
    return p;
}

int64_t part_size(int64_t total, int idx, int parts) { return total > idx ? (total - idx + parts - 1) / parts : 0; }

}  // namespace

gmsb_tc_options normalise_tc_options(const gmsb_tc_options *in) { return normalise(in); }

TcPlan &ensure_plan(Graph &g, const gmsb_tc_options &opt) {
    GMSB_REQUIRE(!g.directed, "triangle kernels need an undirected graph");
    GMSB_REQUIRE(opt.variant >= GMSB_TC_AUTO && opt.variant <= GMSB_TC_BITMAP, "tc: bad variant");
    if (!opt.reuse_plan) { delete g.dag; g.dag = nullptr; }
    if (!g.dag) g.dag = build_degree_dag(g);
    Dag &d = *g.dag;
    if (d.plan && !same_plan(d.plan->opt, opt)) { delete_plan(d.plan); d.plan = nullptr; }
    if (!d.plan) d.plan = build_plan(d, opt);
    return *d.plan;
}

void tc_total(Graph &g, const gmsb_tc_options &opt_in, uint64_t *out, gmsb_tc_stats *stats) {
    GMSB_REQUIRE(!g.directed, "tc_total: graph must be undirected");
    GMSB_REQUIRE(out != nullptr, "tc_total: null output");
    Runtime &r = rt();
    gmsb_tc_options opt = normalise(&opt_in);
    GMSB_REQUIRE(opt.part_index >= 0 && opt.part_index < opt.part_count, "tc_total: bad partition");
    const uint64_t launches0 = r.launches;
    DevTimer t_orient, t_bm, t_mg, t_gl;

    t_orient.start();
    TcPlan &p = ensure_plan(g, opt);
    Dag &d = *g.dag;
    t_orient.stop();

    DevBuf<unsigned long long> total(1);
    total.zero();
    const int P = opt.part_count, pi = opt.part_index;
    const int64_t my_items = part_size(p.n_items, pi, P);
    const int64_t my_merge = part_size(p.n_merge, pi, P);
    const int64_t my_gallop = part_size(p.n_gallop, pi, P);

    t_bm.start();
    if (my_items) {
        GMSB_REQUIRE(p.n_items < (int64_t(1) << 31), "tc: too many bitmap items");
        DevBuf<unsigned int> tickets(2);
        tickets.zero();
        // one persistent wave per window class (the ticket's free is stream-ordered after the kernels)
        auto launch = [&](auto kern, int BLOCK, const Item *items, int64_t cnt, int cap_words, unsigned int *ticket) {
            const int64_t mine = part_size(cnt, pi, P);
            if (mine == 0) return;
            const size_t smem = ((size_t)cap_words + 1) * 4;
            if (smem > 48 * 1024)
                GMSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int resident = 0;
            GMSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, BLOCK, smem));
            GMSB_REQUIRE(resident >= 1, "tc: bitmap kernel does not fit on an SM");
            const int grid = (int)std::min<int64_t>(mine, (int64_t)r.sm_count * resident);
            kern<<<grid, BLOCK, smem, r.stream>>>(items, pi, P, mine, (uint32_t)cap_words, d.off.p, d.nbr.p,
                                                  p.sorted_vals, total.p, ticket);
            launched();
        };
        const int shape = opt.reserved[2];              // experiments: one launch with another CTA shape
        if (shape == 0) {
            // small windows (<= 55 KB of bitmap): four CTAs of 512 threads per SM = 64 resident warps
            launch(k_tc_bitmap<512, 4>, 512, p.items.p, p.n_items_small, p.small_span_words, tickets.p);
            // wide windows: three CTAs per SM
            launch(k_tc_bitmap<512, 3>, 512, p.items.p + p.n_items_small, p.n_items - p.n_items_small,
                   p.max_span_words, tickets.p + 1);
        } else if (shape == 1) launch(k_tc_bitmap<256, 6>, 256, p.items.p, p.n_items, p.max_span_words, tickets.p);
        else if (shape == 2) launch(k_tc_bitmap<1024, 1>, 1024, p.items.p, p.n_items, p.max_span_words, tickets.p);
        else if (shape == 3) launch(k_tc_bitmap<512, 4>, 512, p.items.p, p.n_items, p.max_span_words, tickets.p);
        else if (shape == 4) launch(k_tc_bitmap<256, 8>, 256, p.items.p, p.n_items, p.max_span_words, tickets.p);
        else if (shape == 5) launch(k_tc_bitmap<512, 3, true>, 512, p.items.p, p.n_items, p.max_span_words, tickets.p);
        else launch(k_tc_bitmap<512, 3>, 512, p.items.p, p.n_items, p.max_span_words, tickets.p);
    }
    t_bm.stop();
    t_mg.start();
    if (my_merge) {
        int grid = (int)std::min<int64_t>(ceil_div(my_merge, kMergeWarps), (int64_t)r.sm_count * 16);
        k_tc_merge<<<grid, kMergeWarps * 32, 0, r.stream>>>(p.m_desc.p, p.m_v.p, pi, P, my_merge, d.off.p, d.nbr.p,
                                                            total.p);
        launched();
    }
    t_mg.stop();
    t_gl.start();
    if (my_gallop) {
        int grid = (int)std::min<int64_t>(ceil_div(my_gallop, 8), (int64_t)r.sm_count * 16);
        k_tc_gallop<<<grid, 256, 0, r.stream>>>(p.g_desc.p, p.g_v.p, pi, P, my_gallop, d.off.p, d.nbr.p, total.p);
        launched();
    }
    t_gl.stop();
    *out = total.get(0);

    if (stats) {
        gmsb_tc_stats s{};
        s.triangles = *out;
        s.algorithmic_bytes = p.algorithmic_bytes / P + (pi == 0 ? p.algorithmic_bytes % P : 0);
        s.wedges_checked = p.wedges / P;
        s.oriented_edges = d.m;
        s.edges_bitmap = part_size(p.n_bitmap_edges, pi, P);
        s.edges_merge = my_merge;
        s.edges_gallop = my_gallop;
        s.ms_orient = t_orient.ms();
        s.ms_bitmap = t_bm.ms();
        s.ms_merge = t_mg.ms();
        s.ms_gallop = t_gl.ms();
        s.ms_count = s.ms_bitmap + s.ms_merge + s.ms_gallop;
        s.launches = (int32_t)(r.launches - launches0);
        s.max_dplus = d.max_dplus;
        s.bytes_bitmap = p.bytes_bitmap;
        s.bytes_light = p.bytes_kept - p.bytes_bitmap;
        s.wedges_bitmap = p.wedges_bitmap;
        s.bitmap_items = p.n_items;
        s.bitmap_smem_bytes = (p.max_span_words + 1) * 4;
        *stats = s;
    }
    if (!opt.reuse_plan) { delete g.dag; g.dag = nullptr; }
}

}  // namespace gmsb
