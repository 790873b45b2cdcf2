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

namespace gmsb {

constexpr int kLenBits = 24;                        // descriptor = (start << 24) | len
constexpr uint64_t kLenMask = (1ull << kLenBits) - 1;

struct Item {            // one CTA's share of a hub's incoming descriptors
    int32_t v;
    int32_t count;
    int64_t begin;
};

struct TcPlan {
    gmsb_tc_options opt{};
    int64_t n_desc = 0;                 // descriptors that can close a triangle
    DevBuf<uint64_t> desc;              // grouped by v (ascending), by u inside a group
    DevBuf<uint32_t> desc_v;            // v of each descriptor
    DevBuf<Item> items;                 // bitmap work items
    int64_t n_items = 0;
    int max_span_words = 0;
    DevBuf<uint64_t> m_desc, g_desc;    // light edges for merge / gallop
    DevBuf<vid_t> m_v, g_v;
    int64_t n_merge = 0, n_gallop = 0, n_bitmap_edges = 0;
    uint64_t algorithmic_bytes = 0;     // B_TC over ALL oriented edges
    uint64_t wedges = 0;
    uint64_t bytes_bitmap = 0, bytes_kept = 0, wedges_bitmap = 0;
    // backing stores of the sorted arrays (double buffers keep the result in either half)
    DevBuf<uint32_t> keys_a, keys_b;
    DevBuf<uint64_t> vals_a, vals_b;
    uint32_t *sorted_keys = nullptr;
    uint64_t *sorted_vals = nullptr;
};

void delete_plan(TcPlan *p) { delete p; }

namespace {

// ---- plan construction -----------------------------------------------------------------------------------------
// One warp per vertex u: a descriptor for every out-edge; edges that cannot close a triangle (empty suffix or
// sink v) get the sentinel key n and sort to the tail.
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
            bytes += 4ull * (unsigned long long)((e - b) + dv);
            bool kp = len > 0 && dv > 0;
            keys[s] = kp ? (uint32_t)v : (uint32_t)n;
            vals[s] = ((uint64_t)(s + 1) << kLenBits) | (uint64_t)len;
            if (kp) {
                atomicAdd(&work[v], (unsigned long long)len);
                atomicAdd(&vbytes[v], 4ull * (unsigned long long)((e - b) + dv));
                wedges += len; kept++;
                kbytes += 4ull * (unsigned long long)((e - b) + dv);
            }
        }
    }
    for (int o = 16; o; o >>= 1) {
        bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
        wedges += __shfl_xor_sync(0xffffffffu, wedges, o);
        kept += __shfl_xor_sync(0xffffffffu, kept, o);
        kbytes += __shfl_xor_sync(0xffffffffu, kbytes, o);
    }
    if (lane == 0 && (bytes | kept)) { atomicAdd(&acc[0], bytes); atomicAdd(&acc[1], wedges); atomicAdd(&acc[2], kept); atomicAdd(&acc[3], kbytes); }
}

// inoff[x] = first descriptor whose key >= x  (keys sorted ascending, length cnt)
__global__ void k_lower_bounds(const uint32_t *__restrict__ keys, int64_t cnt, int64_t n, int64_t *__restrict__ inoff) {
    for (int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; x <= n; x += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = cnt;
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if ((int64_t)keys[mid] < x) lo = mid + 1; else hi = mid;
        }
        inoff[x] = lo;
    }
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

__global__ void k_fill_items(int64_t n, const int64_t *__restrict__ inoff, const int64_t *__restrict__ nitems,
                             const int64_t *__restrict__ item_base, Item *__restrict__ items) {
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        int64_t k = nitems[v];
        if (!k) continue;
        int64_t b = inoff[v], cnt = inoff[v + 1] - b;
        int64_t chunk = (cnt + k - 1) / k;
        int64_t w = item_base[v];
        for (int64_t j = 0; j < k; ++j) {
            int64_t s = j * chunk;
            int64_t c = cnt - s < chunk ? cnt - s : chunk;
            Item it;
            it.v = (int32_t)v; it.begin = b + s; it.count = (int32_t)(c > 0 ? c : 0);
            items[w + j] = it;
        }
    }
}

// Light descriptors: flag for merge / gallop by the length ratio of suffix and N+(v).
__global__ void k_flag_light(const uint32_t *__restrict__ keys, const uint64_t *__restrict__ vals, int64_t cnt,
                             const eid_t *__restrict__ off, const int64_t *__restrict__ nitems, int variant,
                             int ratio, uint8_t *__restrict__ fm, uint8_t *__restrict__ fg) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cnt; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t v = keys[i];
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
        if (flag[i]) { odesc[pos[i]] = vals[i]; ov[pos[i]] = (vid_t)keys[i]; }
}

// ---- counting kernels ----------------------------------------------------------------------------------------------
// Hub bitmap kernel.  Persistent CTAs (one wave, grid = SMs x resident CTAs) pull items (v, slice of v's incoming
// suffix descriptors) from a global ticket, heaviest first.  The shared-memory bitmap is zeroed ONCE per CTA; each
// item sets the bits of N+(v), streams its suffixes, then clears exactly the words it set — so the per-item cost
// is O(d+(v)) and independent of how wide the window (v, last(N+(v))] is.  Inside an item the warps pull
// descriptors from a shared-memory ticket (suffix lengths vary by 1000x), load list elements with coalesced
// 4-deep unrolled loads and do ONE branch-free probe per element: out-of-window elements are clamped onto the
// always-zero guard word bm[cap_words].
constexpr int kDescChunk = 4;

__device__ __forceinline__ uint32_t probe(const uint32_t *bm, uint32_t x, uint32_t cap_words) {
    return (bm[min(x >> 5, cap_words)] >> (x & 31)) & 1u;
}

template <int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
k_tc_bitmap(const Item *__restrict__ items, int64_t first, int64_t stride, int64_t count, uint32_t cap_words,
            const eid_t *__restrict__ off, const vid_t *__restrict__ nbr, const uint64_t *__restrict__ desc,
            unsigned long long *__restrict__ total, unsigned int *__restrict__ ticket) {
    extern __shared__ uint32_t bm[];                  // cap_words + 1 words
    __shared__ unsigned long long red[BLOCK / 32];
    __shared__ unsigned int s_item;
    __shared__ int s_next;
    const int tid = threadIdx.x, lane = tid & 31;
    for (uint32_t i = tid; i <= cap_words; i += BLOCK) bm[i] = 0u;
    unsigned long long hits64 = 0;
    for (;;) {
        if (tid == 0) { s_item = atomicAdd(ticket, 1u); s_next = 0; }
        __syncthreads();                              // ticket visible; previous item's clears done
        const int64_t it = (int64_t)s_item;
        if (it >= count) break;
        const Item item = items[first + (count - 1 - it) * stride];        // heaviest (highest v) first
        const vid_t v = item.v;
        const eid_t ob = off[v], oe = off[v + 1];
        const uint32_t base = (uint32_t)v + 1u;
        for (eid_t j = ob + tid; j < oe; j += BLOCK) {
            const uint32_t x = (uint32_t)nbr[j] - base;
            atomicOr(&bm[x >> 5], 1u << (x & 31));
        }
        __syncthreads();                              // bitmap of N+(v) complete

        uint32_t hits = 0;
        const uint64_t *__restrict__ dptr = desc + item.begin;
        const int cnt = item.count;
        for (;;) {
            int d0 = 0;
            if (lane == 0) d0 = atomicAdd(&s_next, kDescChunk);
            d0 = __shfl_sync(0xffffffffu, d0, 0);
            if (d0 >= cnt) break;
            const int nd = min(kDescChunk, cnt - d0);
            const uint64_t mine = lane < nd ? dptr[d0 + lane] : 0ull;
            for (int k = 0; k < nd; ++k) {
                const uint64_t ds = __shfl_sync(0xffffffffu, mine, k);
                const vid_t *__restrict__ p = nbr + (ds >> kLenBits);
                const int len = (int)(ds & kLenMask);
                int j = lane;
                for (; j + 96 < len; j += 128) {      // 4 independent loads in flight per lane
                    const uint32_t x0 = (uint32_t)p[j] - base, x1 = (uint32_t)p[j + 32] - base;
                    const uint32_t x2 = (uint32_t)p[j + 64] - base, x3 = (uint32_t)p[j + 96] - base;
                    hits += probe(bm, x0, cap_words) + probe(bm, x1, cap_words) + probe(bm, x2, cap_words) +
                            probe(bm, x3, cap_words);
                }
                for (; j < len; j += 32) hits += probe(bm, (uint32_t)p[j] - base, cap_words);
            }
        }
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
    if (o.hub_min_work <= 0) o.hub_min_work = 4096;
    return o;
}

bool same_plan(const gmsb_tc_options &a, const gmsb_tc_options &b) {
    return a.variant == b.variant && a.hub_bitmap_bits == b.hub_bitmap_bits && a.gallop_ratio == b.gallop_ratio &&
           a.hub_min_work == b.hub_min_work && a.reserved[0] == b.reserved[0];
}

TcPlan *build_plan(const Dag &d, const gmsb_tc_options &opt) {
    Runtime &r = rt();
    auto *p = new TcPlan();
    try {
        p->opt = opt;
        const int64_t n = d.n, m = d.m;
        GMSB_REQUIRE(d.max_dplus < (1 << kLenBits), "tc: out-degree too large for the descriptor format");
        GMSB_REQUIRE(m < (int64_t(1) << (64 - kLenBits)), "tc: too many edges for the descriptor format");
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
        radix_sort_pairs(p->keys_a.p, p->keys_b.p, p->vals_a.p, p->vals_b.p, m, bits_for((uint64_t)n),
                         &p->sorted_keys, &p->sorted_vals);
        unsigned long long h_acc[4];
        acc.download(h_acc, 4);
        p->bytes_kept = h_acc[3];
        p->algorithmic_bytes = h_acc[0];
        p->wedges = h_acc[1];
        p->n_desc = (int64_t)h_acc[2];
        const int64_t cnt = p->n_desc;
        if (cnt == 0) return p;

        DevBuf<int64_t> inoff(n + 1), nitems(n + 1), item_base(n + 1);
        k_lower_bounds<<<grid_for(n + 1, 256), 256, 0, r.stream>>>(p->sorted_keys, cnt, n, inoff.p); launched();
        DevBuf<int> mxw(1);
        mxw.zero(); nitems.zero();
        PlanParams pp{opt.variant, hub_bits, (long long)opt.hub_min_work,
                      opt.reserved[0] > 0 ? (long long)opt.reserved[0] : 262144ll};
        k_classify<<<grid_for(n, 256), 256, 0, r.stream>>>(d.off.p, d.nbr.p, n, inoff.p, work.p, vbytes.p, cls.p, pp, nitems.p,
                                                          mxw.p);
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
            k_fill_items<<<grid_for(n, 256), 256, 0, r.stream>>>(n, inoff.p, nitems.p, item_base.p, p->items.p);
            launched();
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

void tc_total(Graph &g, const gmsb_tc_options &opt_in, uint64_t *out, gmsb_tc_stats *stats) {
    GMSB_REQUIRE(!g.directed, "tc_total: graph must be undirected");
    GMSB_REQUIRE(out != nullptr, "tc_total: null output");
    Runtime &r = rt();
    gmsb_tc_options opt = normalise(&opt_in);
    GMSB_REQUIRE(opt.part_index >= 0 && opt.part_index < opt.part_count, "tc_total: bad partition");
    GMSB_REQUIRE(opt.variant >= GMSB_TC_AUTO && opt.variant <= GMSB_TC_BITMAP, "tc_total: bad variant");
    const uint64_t launches0 = r.launches;
    DevTimer t_orient, t_bm, t_mg, t_gl;

    t_orient.start();
    if (!opt.reuse_plan) { delete g.dag; g.dag = nullptr; }
    if (!g.dag) g.dag = build_degree_dag(g);
    Dag &d = *g.dag;
    if (d.plan && !same_plan(d.plan->opt, opt)) { delete_plan(d.plan); d.plan = nullptr; }
    if (!d.plan) d.plan = build_plan(d, opt);
    TcPlan &p = *d.plan;
    t_orient.stop();

    DevBuf<unsigned long long> total(1);
    total.zero();
    const int P = opt.part_count, pi = opt.part_index;
    const int64_t my_items = part_size(p.n_items, pi, P);
    const int64_t my_merge = part_size(p.n_merge, pi, P);
    const int64_t my_gallop = part_size(p.n_gallop, pi, P);

    t_bm.start();
    if (my_items) {
        constexpr int BLOCK = 512, MINB = 3;
        auto kern = k_tc_bitmap<BLOCK, MINB>;
        const size_t smem = ((size_t)p.max_span_words + 1) * 4;
        if (smem > 48 * 1024)
            GMSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int resident = 0;
        GMSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, BLOCK, smem));
        GMSB_REQUIRE(resident >= 1, "tc: bitmap kernel does not fit on an SM");
        GMSB_REQUIRE(my_items < (int64_t(1) << 31), "tc: too many bitmap items");
        const int grid = (int)std::min<int64_t>(my_items, (int64_t)r.sm_count * resident);   // one persistent wave
        DevBuf<unsigned int> ticket(1);
        ticket.zero();
        kern<<<grid, BLOCK, smem, r.stream>>>(p.items.p, pi, P, my_items, (uint32_t)p.max_span_words, d.off.p, d.nbr.p,
                                              p.sorted_vals, total.p, ticket.p);
        launched();                                        // (ticket's free is stream-ordered after the kernel)
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
        s.wedges_bitmap = p.wedges_bitmap;
        s.bytes_light = p.bytes_kept - p.bytes_bitmap;
        s.bitmap_items = p.n_items;
        s.bitmap_smem_bytes = p.max_span_words * 4;
        *stats = s;
    }
    if (!opt.reuse_plan) { delete g.dag; g.dag = nullptr; }
}

}  // namespace gmsb
