// sets.cu — device-resident sorted sets: the Set concept of GMS as an object.
//
// Replaces (as a drop-in type for algorithms written against the Set concept):
//   SortedSetBase<int32_t>            gms/representations/sets/sorted_set.h:22-270
//   SortedSetRefBase<int32_t>         gms/representations/sets/sorted_set_ref.h:10-80   (borrowed neighbourhoods)
//   vec_set_{union,intersect,intersect_count,difference}   sets/sorted_set_operations.h:30-106
//
// A set is an ascending, duplicate-free int32 array in HBM — owned, or borrowed from a graph's CSR like SortedSetRef.
// Every binary operation takes SET HANDLES and leaves its result on the device, so a chain such as the
// Bron-Kerbosch step  P' = P ∩ N(v), X' = X ∩ N(v), P = P \ {v}  (maximal_clique_enum/sequential/tomita.h:17-60) or the
// pull step of the approximate degeneracy order (preprocessing/parallel/degeneracy_approx_set.h:60-80) never
// returns to the host.  The "many" forms run one left set against a batch of right sets (or against the
// neighbourhoods of the members of a set) in one launch: one warp per pair, the same kernels as the batched
// neighbourhood API in setops.cu (galloping for skewed pairs, merge path for balanced ones).
#include "common.cuh"
#include "sort.cuh"
#include "isect.cuh"
#include "ops.cuh"
#include "sets.cuh"

#include <vector>

namespace gmsb {

namespace {

constexpr int kWarps = 8;
constexpr int kGallopRatio = 8;

struct Span {
    const vid_t *p;
    int64_t n;
};

__device__ __forceinline__ uint32_t span_count(const vid_t *a, int na, const vid_t *b, int nb, int lane, vid_t *buf) {
    const int lo = na < nb ? na : nb, hi = na < nb ? nb : na;
    if (lo == 0) return 0;
    if ((long long)hi >= (long long)kGallopRatio * lo) return warp_gallop_count(a, na, b, nb, lane);
    return warp_merge_count(a, na, b, nb, lane, buf);
}

// out[i] = |A ∩ B_i| (op 0) or |A ∪ B_i| (op 1) or |A \ B_i| (op 2); `left` is one span shared by all pairs
__global__ void __launch_bounds__(kWarps * 32)
k_span_count(Span left, const Span *__restrict__ right, int64_t np, int op, unsigned long long *__restrict__ out) {
    __shared__ vid_t stage[kWarps][kMergeTile + 2];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < np; i += nwarps) {
        const Span b = right[i];
        unsigned long long c = span_count(left.p, (int)left.n, b.p, (int)b.n, lane, stage[wib]);
        c = warp_sum(c);
        if (lane == 0)
            out[i] = op == 0 ? c : (op == 1 ? (unsigned long long)left.n + (unsigned long long)b.n - c
                                            : (unsigned long long)left.n - c);
    }
}
// neighbourhoods of the listed vertices as the right-hand spans
__global__ void k_neighbourhood_spans(const eid_t *__restrict__ off, const vid_t *__restrict__ nbr, int64_t n,
                                      const vid_t *__restrict__ verts, int64_t nv, Span *__restrict__ out,
                                      int *__restrict__ bad) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
        const vid_t v = verts[i];
        if (v < 0 || v >= n) { *bad = 1; out[i] = Span{nullptr, 0}; continue; }
        out[i] = Span{nbr + off[v], off[v + 1] - off[v]};
    }
}

// Materialising A op B_i, ascending, two passes (count, then write at out_off[i]); one warp per pair.
//   op 0 intersect: elements of the shorter list found in the longer one      (std::set_intersection)
//   op 2 difference: elements of A not found in B                              (std::set_difference)
//   op 1 union: every element computes its own slot — j + lower_bound(other, x) - common elements before it
template <int OP>
__global__ void __launch_bounds__(256)
k_span_op(Span left, const Span *__restrict__ right, int64_t np, int64_t *__restrict__ counts,
          const int64_t *__restrict__ out_off, vid_t *__restrict__ out_elems) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < np; i += nwarps) {
        const vid_t *a = left.p, *b = right[i].p;
        int na = (int)left.n, nb = (int)right[i].n;
        if (OP == 0 && na > nb) { const vid_t *t = a; a = b; b = t; const int tn = na; na = nb; nb = tn; }
        const int64_t w = out_elems ? out_off[i] : 0;
        int64_t c = 0;
        int common = 0;
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
            const bool keep = OP == 0 ? hit : (OP == 2 ? (j < na && !hit) : j < na);
            const unsigned hmask = __ballot_sync(0xffffffffu, hit), kmask = __ballot_sync(0xffffffffu, keep);
            if (out_elems && keep) {
                if (OP == 1) out_elems[w + j + lo - (common + __popc(hmask & ((1u << lane) - 1)))] = x;
                else out_elems[w + c + __popc(kmask & ((1u << lane) - 1))] = x;
            }
            c += __popc(kmask);
            common += __popc(hmask);
        }
        if (OP == 1) {
            if (!out_elems) { if (lane == 0) counts[i] = (int64_t)na + nb - common; continue; }
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
                const unsigned hmask = __ballot_sync(0xffffffffu, hit);
                if (t < nb && !hit) out_elems[w + t + lo - (common + __popc(hmask & ((1u << lane) - 1)))] = x;
                common += __popc(hmask);
            }
        } else if (!out_elems && lane == 0) {
            counts[i] = c;
        }
    }
}

__global__ void k_iota(vid_t *__restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (vid_t)i;
}
// flags[0] = 1 when x is in the set; flags[1] = lower bound position of x
__global__ void k_find(const vid_t *__restrict__ p, int64_t n, vid_t x, int64_t *__restrict__ res) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (p[mid] < x) lo = mid + 1; else hi = mid; }
    res[0] = lo < n && p[lo] == x;
    res[1] = lo;
}
__global__ void k_equal(const vid_t *__restrict__ a, const vid_t *__restrict__ b, int64_t n, int *__restrict__ diff) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (a[i] != b[i]) *diff = 1;
}
__global__ void k_insert(const vid_t *__restrict__ in, int64_t n, int64_t pos, vid_t x, vid_t *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = i < pos ? in[i] : (i == pos ? x : in[i - 1]);
}
__global__ void k_erase(const vid_t *__restrict__ in, int64_t n, int64_t pos, vid_t *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i + 1 < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = i < pos ? in[i] : in[i + 1];
}

Span span_of(const DevSet &s) { return Span{s.p, s.n}; }

template <typename K>
void launch_op(int op, K k0, K k1, K k2, int grid, Span left, const Span *right, int64_t np, int64_t *counts,
               const int64_t *out_off, vid_t *out_elems) {
    Runtime &r = rt();
    (op == 0 ? k0 : (op == 1 ? k1 : k2))<<<grid, 256, 0, r.stream>>>(left, right, np, counts, out_off, out_elems);
    launched();
}

}  // namespace

DevSet *set_from_device(const vid_t *src, int64_t count, bool sorted) {
    GMSB_REQUIRE(count >= 0 && (count == 0 || src != nullptr), "set: bad arguments");
    GMSB_REQUIRE(count < (int64_t(1) << 31), "set: more than 2^31 elements");
    Runtime &r = rt();
    auto *s = new DevSet();
    try {
        s->own.alloc((size_t)count);
        s->n = count;
        if (count) {
            if (sorted) {
                GMSB_CUDA(cudaMemcpyAsync(s->own.p, src, sizeof(vid_t) * count, cudaMemcpyDeviceToDevice, r.stream));
            } else {
                // SortedSetBase(const T*, size_t) sorts what it is given (sorted_set.h:64-66,265-268)
                DevBuf<vid_t> tmp((size_t)count);
                GMSB_CUDA(cudaMemcpyAsync(tmp.p, src, sizeof(vid_t) * count, cudaMemcpyDeviceToDevice, r.stream));
                size_t bytes = 0;
                GMSB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, tmp.p, s->own.p, (int)count, 0, 32, r.stream));
                DevBuf<uint8_t> scratch(bytes);
                GMSB_CUDA(cub::DeviceRadixSort::SortKeys(scratch.p, bytes, tmp.p, s->own.p, (int)count, 0, 32, r.stream));
                r.launches += 5;
                GMSB_CUDA(cudaStreamSynchronize(r.stream));
            }
        }
        s->p = s->own.p;
        GMSB_CUDA(cudaStreamSynchronize(r.stream));
    } catch (...) { delete s; throw; }
    return s;
}

DevSet *set_from_host(const vid_t *elems, int64_t count) {
    GMSB_REQUIRE(count >= 0 && (count == 0 || elems != nullptr), "set_from_host: bad arguments");
    for (int64_t i = 0; i < count; ++i) GMSB_REQUIRE(elems[i] >= 0, "set_from_host: negative element");
    bool sorted = true;
    for (int64_t i = 1; i < count && sorted; ++i) sorted = elems[i - 1] < elems[i];
    DevBuf<vid_t> tmp((size_t)count);
    tmp.upload(elems, (size_t)count);
    return set_from_device(tmp.p, count, sorted);
}

DevSet *set_range(int64_t bound) {
    GMSB_REQUIRE(bound >= 0 && bound < (int64_t(1) << 31), "set_range: bad bound");
    Runtime &r = rt();
    auto *s = new DevSet();
    try {
        s->own.alloc((size_t)bound);
        s->n = bound;
        s->p = s->own.p;
        if (bound) { k_iota<<<grid_for(bound, 256), 256, 0, r.stream>>>(s->own.p, bound); launched(); }
        GMSB_CUDA(cudaStreamSynchronize(r.stream));
    } catch (...) { delete s; throw; }
    return s;
}

DevSet *set_neighbourhood(Graph &g, vid_t v) {
    GMSB_REQUIRE(v >= 0 && v < g.n, "set_neighbourhood: vertex out of range");
    eid_t be[2];
    GMSB_CUDA(cudaMemcpyAsync(be, g.off.p + v, 2 * sizeof(eid_t), cudaMemcpyDeviceToHost, rt().stream));
    GMSB_CUDA(cudaStreamSynchronize(rt().stream));
    auto *s = new DevSet();
    s->p = g.nbr.p + be[0];
    s->n = be[1] - be[0];
    return s;
}

DevSet *set_clone(const DevSet &a) { return set_from_device(a.p, a.n, true); }

void set_to_host(const DevSet &a, vid_t *out) {
    if (a.n == 0) return;
    GMSB_REQUIRE(out != nullptr, "set_to_host: null output");
    GMSB_CUDA(cudaMemcpyAsync(out, a.p, sizeof(vid_t) * a.n, cudaMemcpyDeviceToHost, rt().stream));
    GMSB_CUDA(cudaStreamSynchronize(rt().stream));
}

static void find_in(const DevSet &a, vid_t x, int64_t res[2]) {
    Runtime &r = rt();
    DevBuf<int64_t> d(2);
    k_find<<<1, 1, 0, r.stream>>>(a.p, a.n, x, d.p); launched();
    d.download(res, 2);
}
bool set_contains(const DevSet &a, vid_t x) {
    if (a.n == 0) return false;
    int64_t res[2];
    find_in(a, x, res);
    return res[0] != 0;
}
bool set_equal(const DevSet &a, const DevSet &b) {
    if (a.n != b.n) return false;
    if (a.n == 0 || a.p == b.p) return true;
    Runtime &r = rt();
    DevBuf<int> diff(1);
    diff.zero();
    k_equal<<<grid_for(a.n, 256), 256, 0, r.stream>>>(a.p, b.p, a.n, diff.p); launched();
    return diff.get(0) == 0;
}
// add / remove keep the set sorted (sorted_set.h:222-243); a borrowed view becomes an owning set first
void set_add(DevSet &a, vid_t x) {
    GMSB_REQUIRE(x >= 0, "set_add: negative element");
    Runtime &r = rt();
    int64_t res[2] = {0, 0};
    if (a.n) find_in(a, x, res);
    if (res[0]) return;
    DevBuf<vid_t> grown((size_t)a.n + 1);
    k_insert<<<grid_for(a.n + 1, 256), 256, 0, r.stream>>>(a.p, a.n, res[1], x, grown.p); launched();
    GMSB_CUDA(cudaStreamSynchronize(r.stream));
    a.own = std::move(grown);
    a.p = a.own.p;
    a.n += 1;
}
void set_remove(DevSet &a, vid_t x) {
    if (a.n == 0) return;
    Runtime &r = rt();
    int64_t res[2];
    find_in(a, x, res);
    if (!res[0]) return;
    DevBuf<vid_t> shrunk((size_t)a.n - 1);
    if (a.n > 1) { k_erase<<<grid_for(a.n, 256), 256, 0, r.stream>>>(a.p, a.n, res[1], shrunk.p); launched(); }
    GMSB_CUDA(cudaStreamSynchronize(r.stream));
    a.own = std::move(shrunk);
    a.p = a.own.p;
    a.n -= 1;
}

// counts[i] for one left set against np right spans (device array)
static void count_spans(const DevSet &a, const Span *right_dev, int64_t np, int op, uint64_t *out_host) {
    Runtime &r = rt();
    DevBuf<unsigned long long> out((size_t)np);
    const int grid = (int)std::min<int64_t>(ceil_div(np, kWarps), (int64_t)r.sm_count * 16);
    k_span_count<<<grid, kWarps * 32, 0, r.stream>>>(span_of(a), right_dev, np, op, out.p); launched();
    out.download(reinterpret_cast<unsigned long long *>(out_host), (size_t)np);
}

void set_op_count_many(int op, const DevSet &a, int64_t np, DevSet *const *bs, uint64_t *out) {
    GMSB_REQUIRE(op >= 0 && op <= 2, "set op: unknown operation");
    if (np == 0) return;
    std::vector<Span> h((size_t)np);
    for (int64_t i = 0; i < np; ++i) { GMSB_REQUIRE(bs[i] != nullptr, "set op: null set handle"); h[i] = span_of(*bs[i]); }
    DevBuf<Span> d((size_t)np);
    d.upload(h.data(), (size_t)np);
    count_spans(a, d.p, np, op, out);
}

void set_op_count_members(int op, const DevSet &a, Graph &g, const DevSet &members, uint64_t *out) {
    GMSB_REQUIRE(op >= 0 && op <= 2, "set op: unknown operation");
    if (members.n == 0) return;
    Runtime &r = rt();
    DevBuf<Span> d((size_t)members.n);
    DevBuf<int> bad(1);
    bad.zero();
    k_neighbourhood_spans<<<grid_for(members.n, 256), 256, 0, r.stream>>>(g.off.p, g.nbr.p, g.n, members.p, members.n,
                                                                        d.p, bad.p);
    launched();
    GMSB_REQUIRE(bad.get(0) == 0, "set op: member outside the graph");
    count_spans(a, d.p, members.n, op, out);
}

void set_op_many(int op, const DevSet &a, int64_t np, DevSet *const *bs, DevSet **outs) {
    GMSB_REQUIRE(op >= 0 && op <= 2, "set op: unknown operation");
    if (np == 0) return;
    Runtime &r = rt();
    std::vector<Span> h((size_t)np);
    for (int64_t i = 0; i < np; ++i) { GMSB_REQUIRE(bs[i] != nullptr, "set op: null set handle"); h[i] = span_of(*bs[i]); }
    DevBuf<Span> d((size_t)np);
    d.upload(h.data(), (size_t)np);
    DevBuf<int64_t> counts((size_t)np + 1), offs((size_t)np + 1);
    counts.zero();
    const int grid = (int)std::min<int64_t>(ceil_div(np, 8), (int64_t)r.sm_count * 16);
    launch_op(op, k_span_op<0>, k_span_op<1>, k_span_op<2>, grid, span_of(a), d.p, np, counts.p, nullptr, nullptr);
    exclusive_sum(counts.p, offs.p, np + 1);
    std::vector<int64_t> ho((size_t)np + 1);
    offs.download(ho.data(), (size_t)np + 1);
    DevBuf<vid_t> elems((size_t)ho[np]);
    if (ho[np])
        launch_op(op, k_span_op<0>, k_span_op<1>, k_span_op<2>, grid, span_of(a), d.p, np, nullptr, offs.p, elems.p);
    std::vector<DevSet *> made;
    try {
        for (int64_t i = 0; i < np; ++i) {
            made.push_back(set_from_device(elems.p + ho[i], ho[i + 1] - ho[i], true));
            outs[i] = made.back();
        }
    } catch (...) { for (auto *m : made) delete m; throw; }
}

}  // namespace gmsb
