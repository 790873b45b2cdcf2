// graph_build.cu — device-resident CSR construction.
//
// Replaces, on the GPU:
//   BuilderBase::MakeGraphFromEL + SquishGraph   gms/third_party/gapbs/builder.h:279-298,206-251
//   BuilderBase::RelabelByDegree                 gms/third_party/gapbs/builder.h:1699-1735
//   SetGraph<Set>::FromCGraph                    gms/representations/graphs/set_graph.h:153-181
//
// Design: the reference scatters edges with fetch_and_add and then sorts every list on the CPU.  Here the whole
// edge multiset is one array of 64-bit keys (u<<32 | v); a single LSD radix sort over only the significant bits
// puts it in CSR order, duplicates and self loops are dropped by a flag + scan compaction, and the offsets
// fall out of a degree histogram + scan.  All passes are streaming and HBM-bound.
#include "common.cuh"
#include "sort.cuh"
#include "isect.cuh"
#include "orient.cuh"
#include <vector>

namespace gmsb {

namespace {

__global__ void k_max_id(const vid_t *__restrict__ a, const vid_t *__restrict__ b, int64_t m, int *out) {
    int mx = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
        mx = max(mx, max(a[i], b[i]));
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, mx);
}

__global__ void k_min_id(const vid_t *__restrict__ a, const vid_t *__restrict__ b, int64_t m, int *out) {
    int mn = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
        mn = min(mn, min(a[i], b[i]));
    for (int o = 16; o; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    if ((threadIdx.x & 31) == 0) atomicMin(out, mn);
}

__global__ void k_make_keys(const vid_t *__restrict__ src, const vid_t *__restrict__ dst, int64_t m, bool sym,
                            uint64_t *__restrict__ keys) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t u = (uint32_t)src[i], v = (uint32_t)dst[i];
        keys[i] = (u << 32) | v;
        if (sym) keys[m + i] = (v << 32) | u;
    }
}

// keep[i] = 1 for the first copy of each (u,v) with u != v; also histogram kept entries per source vertex.
__global__ void k_flag_unique(const uint64_t *__restrict__ keys, int64_t K, uint8_t *__restrict__ keep,
                              unsigned long long *__restrict__ deg /* n+1, counts land at [u+1] */) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < K; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t k = keys[i];
        uint32_t u = (uint32_t)(k >> 32), v = (uint32_t)k;
        bool kp = (u != v) && (i == 0 || keys[i - 1] != k);
        keep[i] = kp;
        if (kp) atomicAdd(&deg[u + 1], 1ull);
    }
}

__global__ void k_scatter_kept(const uint64_t *__restrict__ keys, const uint8_t *__restrict__ keep,
                               const int64_t *__restrict__ pos, int64_t K, vid_t *__restrict__ nbr) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < K; i += (int64_t)gridDim.x * blockDim.x)
        if (keep[i]) nbr[pos[i]] = (vid_t)(uint32_t)keys[i];
}

__global__ void k_degree_keys_desc(const eid_t *__restrict__ off, int64_t n, uint64_t *__restrict__ keys) {
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x)
        keys[v] = ((uint64_t)(off[v + 1] - off[v]) << 32) | (uint32_t)v;
}

// position i of the (degree desc, id desc) order holds old vertex key&0xffffffff: record new id and new degree.
__global__ void k_newid_from_order(const uint64_t *__restrict__ sorted, int64_t n, vid_t *__restrict__ newid,
                                   eid_t *__restrict__ newoff /* n+1 */) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t k = sorted[i];
        newid[(uint32_t)k] = (vid_t)i;
        newoff[i + 1] = (eid_t)(k >> 32);
        if (i == 0) newoff[0] = 0;
    }
}

// one warp per vertex: emit (newid[u]<<32 | newid[v]) for every slot, in place of the old slot.
__global__ void k_relabel_keys(const eid_t *__restrict__ off, const vid_t *__restrict__ nbr, int64_t n,
                               const vid_t *__restrict__ newid, uint64_t *__restrict__ keys) {
    int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warp; u < n; u += nwarps) {
        eid_t b = off[u], e = off[u + 1];
        uint64_t hi = (uint64_t)(uint32_t)newid[u] << 32;
        for (eid_t s = b + lane; s < e; s += 32) keys[s] = hi | (uint32_t)newid[nbr[s]];
    }
}

// Validation of an untrusted CSR in two streaming passes (offsets first: the slot pass trusts them).
// Per vertex: offsets monotone and inside [0, slots] (flags[0] otherwise); counts the descents nbr[b-1] > nbr[b] that
// sit on the first slot b of a non-empty list — those are legitimate.
__global__ void k_check_offsets(const eid_t *__restrict__ off, const vid_t *__restrict__ nbr, int64_t n, int64_t slots,
                                int *__restrict__ flags, unsigned long long *__restrict__ boundary_descents) {
    unsigned long long c = 0;
    for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < n; u += (int64_t)gridDim.x * blockDim.x) {
        const eid_t b = off[u], e = off[u + 1];
        if (b < 0 || e < b || e > slots) { flags[0] = 1; continue; }
        if (e > b && b > 0 && nbr[b - 1] > nbr[b]) ++c;
    }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(boundary_descents, c);
}
// offsets only (the upload pipeline validates them before the neighbour slots have arrived)
__global__ void k_check_offsets_only(const eid_t *__restrict__ off, int64_t n, int64_t slots, int *__restrict__ flags) {
    for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < n; u += (int64_t)gridDim.x * blockDim.x) {
        const eid_t b = off[u], e = off[u + 1];
        if (b < 0 || e < b || e > slots) flags[0] = 1;
    }
}
// Per slot: id inside [0, n) (flags[0] otherwise); counts every descent nbr[s-1] > nbr[s].  Some list is unsorted
// exactly when there are more descents than list boundaries that explain them.
__global__ void k_check_slots(const vid_t *__restrict__ nbr, int64_t first, int64_t slots, int64_t n,
                              int *__restrict__ flags, unsigned long long *__restrict__ descents) {
    unsigned long long c = 0;
    for (int64_t s = first + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < slots;
         s += (int64_t)gridDim.x * blockDim.x) {
        const vid_t v = nbr[s];
        if (v < 0 || v >= n) flags[0] = 1;
        if (s > 0 && nbr[s - 1] > v) ++c;
    }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(descents, c);
}
// (u << 32 | v) for every slot: a radix sort of these keys sorts each list and keeps the lists in place
__global__ void k_slot_keys(const eid_t *__restrict__ off, const vid_t *__restrict__ nbr, int64_t n,
                            uint64_t *__restrict__ keys) {
    int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warp; u < n; u += nwarps)
        for (eid_t s = off[u] + lane; s < off[u + 1]; s += 32) keys[s] = ((uint64_t)u << 32) | (uint32_t)nbr[s];
}

__global__ void k_low32(const uint64_t *__restrict__ keys, int64_t K, vid_t *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < K; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (vid_t)(uint32_t)keys[i];
}


}  // namespace

Graph *graph_from_csr_device(int64_t n, const eid_t *off, const vid_t *nbr, bool directed, bool host_src) {
    GMSB_REQUIRE(n >= 0 && off != nullptr, "graph_from_csr: bad arguments");
    Runtime &r = rt();
    auto *g = new Graph();
    try {
        g->n = n;
        g->directed = directed;
        g->off.alloc(n + 1);
        cudaMemcpyKind kind = host_src ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
        GMSB_CUDA(cudaMemcpyAsync(g->off.p, off, sizeof(eid_t) * (n + 1), kind, r.stream));
        eid_t first, last;
        if (host_src) { first = off[0]; last = off[n]; }
        else {
            GMSB_CUDA(cudaMemcpyAsync(&first, off, sizeof(eid_t), cudaMemcpyDeviceToHost, r.stream));
            GMSB_CUDA(cudaMemcpyAsync(&last, off + n, sizeof(eid_t), cudaMemcpyDeviceToHost, r.stream));
            GMSB_CUDA(cudaStreamSynchronize(r.stream));
        }
        GMSB_REQUIRE(first == 0 && last >= 0, "graph_from_csr: offsets must start at 0 and be non-negative");
        GMSB_REQUIRE(last == 0 || nbr != nullptr, "graph_from_csr: null neighbour array");
        g->slots = last;
        g->nbr.alloc(last);
        if (last) GMSB_CUDA(cudaMemcpyAsync(g->nbr.p, nbr, sizeof(vid_t) * last, kind, r.stream));
        // SortedSet's constructor sorts what it is given (sorted_set.h:64-66, so FromCGraph accepts any list order);
        // here one streaming pass checks monotone offsets, id range and order, and only an unsorted CSR pays for a sort.
        if (last && n) {
            DevBuf<int> flags(1);
            DevBuf<unsigned long long> desc(2);
            flags.zero(); desc.zero();
            k_check_offsets<<<grid_for(n, 256), 256, 0, r.stream>>>(g->off.p, g->nbr.p, n, last, flags.p, desc.p);
            launched();
            GMSB_REQUIRE(flags.get(0) == 0, "graph_from_csr: offsets not monotone or outside the neighbour array");
            k_check_slots<<<grid_for(last, 256), 256, 0, r.stream>>>(g->nbr.p, 0, last, n, flags.p, desc.p + 1); launched();
            unsigned long long h[2];
            desc.download(h, 2);
            GMSB_REQUIRE(flags.get(0) == 0, "graph_from_csr: neighbour id out of range");
            if (h[1] > h[0]) {
                DevBuf<uint64_t> keys(last), alt(last);
                k_slot_keys<<<grid_for(n * 32, 256), 256, 0, r.stream>>>(g->off.p, g->nbr.p, n, keys.p); launched();
                uint64_t *sorted = radix_sort_keys(keys.p, alt.p, last, 0, 32 + bits_for((uint64_t)(n - 1)));
                k_low32<<<grid_for(last, 256), 256, 0, r.stream>>>(sorted, last, g->nbr.p); launched();
            }
        }
        GMSB_CUDA(cudaStreamSynchronize(r.stream));
    } catch (...) { delete g; throw; }
    return g;
}

// gmsb_graph_from_csr_ex(GMSB_BUILD_ORIENT): the host CSR is uploaded in chunks on a copy stream while the library
// stream already works on what has arrived — the degree ranking needs the offsets only (the first 6 % of the bytes);
// the validation and the relabel / emit / sort passes of the orientation run per vertex range as soon as that range's
// neighbour slots are on the device.  When the last chunk lands only that chunk's passes and the move of the finished
// rows into rank order are left.  (Host memory should be pinned; with pageable memory the copies are staged by the
// driver and overlap less.)
namespace {
struct UploadStreams {
    cudaStream_t copy = nullptr;
    std::vector<cudaEvent_t> ev;
    ~UploadStreams() {
        for (auto e : ev) cudaEventDestroy(e);
        if (copy) cudaStreamDestroy(copy);
    }
    cudaEvent_t mark() {
        cudaEvent_t e;
        GMSB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ev.push_back(e);
        GMSB_CUDA(cudaEventRecord(e, copy));
        return e;
    }
};

// Uploads the offsets and the neighbour slots of the vertex range [u0, u1) (the whole graph, or one device's share of
// it) in `chunks` pieces of about equal slot counts and runs the orientation passes on every piece as it lands.
// flags[0] is raised by an id outside [0, n); desc[1] counts the descents between consecutive slots.
// off_dev (optional): the complete offsets array already on this device (ordered on the library stream) — then only the
// few entries the host needs for the cuts are read from `off`.
void upload_and_orient_range(Graph *g, OrientPipeline &pipe, UploadStreams &st, const eid_t *off, const vid_t *nbr,
                             int64_t u0, int64_t u1, int chunks, DevBuf<int> &flags, DevBuf<unsigned long long> &desc,
                             const eid_t *off_dev = nullptr) {
    Runtime &r = rt();
    const int64_t n = g->n;
    const eid_t last = g->slots;
    GMSB_CUDA(cudaStreamCreateWithFlags(&st.copy, cudaStreamNonBlocking));
    // allocations above may have been served from blocks whose last use is still queued on the library stream
    GMSB_CUDA(cudaStreamSynchronize(r.stream));
    cudaEvent_t ev_off = nullptr;
    if (off_dev) {
        GMSB_CUDA(cudaMemcpyAsync(g->off.p, off_dev, sizeof(eid_t) * (n + 1), cudaMemcpyDeviceToDevice, r.stream));
    } else {
        GMSB_CUDA(cudaMemcpyAsync(g->off.p, off, sizeof(eid_t) * (n + 1), cudaMemcpyHostToDevice, st.copy));
        ev_off = st.mark();
    }
    // vertex ranges of about equal slot counts (the host array is only used to pick the cuts: a malformed one is
    // caught by the device check of the offsets below before any range is processed)
    auto clamp = [&](eid_t x) { return x < 0 ? eid_t(0) : (x > last ? last : x); };
    const eid_t lo_slot = clamp(off[u0]), hi_slot = std::max(lo_slot, clamp(off[u1]));
    std::vector<int64_t> cut(1, u0);
    for (int c = 1; c < chunks; ++c) {
        const eid_t target = lo_slot + (hi_slot - lo_slot) / chunks * c;
        int64_t lo = cut.back(), hi = u1;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (off[mid] < target) lo = mid + 1; else hi = mid; }
        cut.push_back(lo);
    }
    cut.push_back(u1);
    std::vector<cudaEvent_t> ev_chunk;
    std::vector<eid_t> upto_slot;
    eid_t sent = lo_slot;
    for (int c = 0; c < chunks; ++c) {
        eid_t upto = c + 1 == chunks ? hi_slot : off[cut[c + 1]];
        if (upto < sent || upto > hi_slot) upto = sent;                     // malformed offsets: rejected below
        if (upto > sent)
            GMSB_CUDA(cudaMemcpyAsync(g->nbr.p + sent, nbr + sent, sizeof(vid_t) * (size_t)(upto - sent),
                                      cudaMemcpyHostToDevice, st.copy));
        sent = upto;
        upto_slot.push_back(upto);
        ev_chunk.push_back(st.mark());
    }
    // library stream: offsets -> validation -> ranking; then range by range
    if (ev_off) GMSB_CUDA(cudaStreamWaitEvent(r.stream, ev_off, 0));
    k_check_offsets_only<<<grid_for(n, 256), 256, 0, r.stream>>>(g->off.p, n, last, flags.p); launched();
    GMSB_REQUIRE(flags.get(0) == 0, "graph_from_csr: offsets not monotone or outside the neighbour array");
    orient_pipeline_begin(*g, pipe);
    eid_t checked = lo_slot;
    for (int c = 0; c < chunks; ++c) {
        GMSB_CUDA(cudaStreamWaitEvent(r.stream, ev_chunk[c], 0));
        const eid_t upto = upto_slot[c];
        if (upto > checked) {
            k_check_slots<<<grid_for(upto - checked, 256), 256, 0, r.stream>>>(g->nbr.p, checked, upto, n, flags.p,
                                                                              desc.p + 1);
            launched();
            checked = upto;
        }
        orient_pipeline_range(*g, pipe, cut[c], cut[c + 1]);
    }
}
}  // namespace

Graph *graph_from_csr_host_pipelined(int64_t n, const eid_t *off, const vid_t *nbr) {
    GMSB_REQUIRE(n >= 0 && off != nullptr, "graph_from_csr: bad arguments");
    Runtime &r = rt();
    const eid_t last = off[n];
    GMSB_REQUIRE(off[0] == 0 && last >= 0, "graph_from_csr: offsets must start at 0 and be non-negative");
    GMSB_REQUIRE(last == 0 || nbr != nullptr, "graph_from_csr: null neighbour array");
    if (n == 0 || last == 0) return graph_from_csr_device(n, off, nbr, false, true);
    UploadStreams st;
    auto *g = new Graph();
    OrientPipeline pipe;
    try {
        g->n = n; g->slots = last; g->directed = false;
        g->off.alloc(n + 1);
        g->nbr.alloc(last);
        DevBuf<int> flags(1);
        DevBuf<unsigned long long> desc(2);
        flags.zero(); desc.zero();
        upload_and_orient_range(g, pipe, st, off, nbr, 0, n, 16, flags, desc);
        k_check_offsets<<<grid_for(n, 256), 256, 0, r.stream>>>(g->off.p, g->nbr.p, n, last, flags.p, desc.p); launched();
        unsigned long long h[2];
        desc.download(h, 2);
        GMSB_REQUIRE(flags.get(0) == 0, "graph_from_csr: neighbour id out of range");
        orient_pipeline_finish(*g, pipe);
        g->dag = pipe.d;
        pipe.d = nullptr;
        g->dag_pinned = true;
        if (h[1] > h[0]) {          // some list was not ascending: sort the lists (the DAG does not depend on their order)
            DevBuf<uint64_t> keys(last), alt(last);
            k_slot_keys<<<grid_for(n * 32, 256), 256, 0, r.stream>>>(g->off.p, g->nbr.p, n, keys.p); launched();
            uint64_t *sorted = radix_sort_keys(keys.p, alt.p, last, 0, 32 + bits_for((uint64_t)(n - 1)));
            k_low32<<<grid_for(last, 256), 256, 0, r.stream>>>(sorted, last, g->nbr.p); launched();
        }
        GMSB_CUDA(cudaStreamSynchronize(r.stream));
    } catch (...) {
        cudaStreamSynchronize(st.copy);
        cudaStreamSynchronize(r.stream);
        delete pipe.d;
        delete g;
        throw;
    }
    return g;
}

// ---- sharded construction (one process per device) ----------------------------------------------------------------------
// Device `part` of `parts` uploads the offsets and the neighbour slots of ITS vertex range only (ranges of about equal
// slot counts, computed the same way by everyone from the host offsets), orients and sorts those rows, and packs them
// into a piece.  The offsets are needed in full (the ranking looks at every degree): either every device uploads them
// (off_dev == nullptr), or the caller uploads 1/parts of them per device, all-gathers and passes the device copy.  The caller exchanges the pieces (one all-gather) and the d+ values (one all-reduce of an int32[n] that
// every device filled in its own range); shard_finish then moves all rows into rank order.  Per device the host link
// carries 1/parts of the neighbour array and the orientation passes touch 1/parts of the rows.  The resulting handle
// holds the oriented representation only (Graph::dag_only): the symmetric lists of the other ranges never arrive.
Shard::~Shard() { delete d; delete g; }

Shard *shard_begin(int64_t n, const eid_t *off, const vid_t *nbr, const eid_t *off_dev, int part, int parts,
                   int64_t *piece_len) {
    GMSB_REQUIRE(n >= 0 && off != nullptr && piece_len != nullptr, "shard_begin: bad arguments");
    GMSB_REQUIRE(parts >= 1 && parts <= 64 && part >= 0 && part < parts, "shard_begin: bad partition");
    Runtime &r = rt();
    const eid_t last = off[n];
    GMSB_REQUIRE(off[0] == 0 && last >= 0, "graph_from_csr: offsets must start at 0 and be non-negative");
    GMSB_REQUIRE(last == 0 || nbr != nullptr, "graph_from_csr: null neighbour array");
    auto *s = new Shard();
    UploadStreams st;
    OrientPipeline pipe;
    try {
        s->part = part; s->parts = parts;
        s->cut.assign((size_t)parts + 1, n);
        s->cut[0] = 0;
        for (int i = 1; i < parts; ++i) {           // first vertex whose list starts at or after slot last / parts * i
            const eid_t target = last / parts * i;
            int64_t lo = s->cut[(size_t)i - 1], hi = n;
            while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (off[mid] < target) lo = mid + 1; else hi = mid; }
            s->cut[(size_t)i] = lo;
        }
        const int64_t u0 = s->cut[(size_t)part], u1 = s->cut[(size_t)part + 1];
        Graph *g = s->g = new Graph();
        g->n = n; g->slots = last; g->directed = false; g->dag_only = true;
        g->off.alloc(n + 1);
        g->nbr.alloc(last);                 // only the slots of [u0, u1) are ever filled
        DevBuf<int> flags(1);
        DevBuf<unsigned long long> desc(2);
        flags.zero(); desc.zero();
        upload_and_orient_range(g, pipe, st, off, nbr, u0, u1, 4, flags, desc, off_dev);
        s->d = pipe.d;
        pipe.d = nullptr;
        GMSB_REQUIRE(flags.get(0) == 0, "graph_from_csr: neighbour id out of range");
        s->piece_len = orient_piece_layout(*g, s->d->rank.p, pipe.w, u0, u1, s->piece_off);
        s->w = std::move(pipe.w);
        *piece_len = s->piece_len;
        GMSB_CUDA(cudaStreamSynchronize(r.stream));
    } catch (...) {
        cudaStreamSynchronize(st.copy);
        cudaStreamSynchronize(r.stream);
        delete pipe.d;
        delete s;
        throw;
    }
    return s;
}

void shard_export(Shard &s, vid_t *piece_dev, int32_t *dplus_all_dev) {
    GMSB_REQUIRE(dplus_all_dev != nullptr && (piece_dev != nullptr || s.piece_len == 0), "shard_export: null buffer");
    GMSB_REQUIRE(s.g != nullptr && s.d != nullptr, "shard_export: the shard has been finished");
    orient_piece_export(*s.g, s.w, s.cut[(size_t)s.part], s.cut[(size_t)s.part + 1], s.piece_off, piece_dev, dplus_all_dev);
}

Graph *shard_finish(Shard &s, const vid_t *pieces_dev, int64_t piece_stride, const int32_t *dplus_all_dev) {
    GMSB_REQUIRE(s.g != nullptr && s.d != nullptr, "shard_finish: the shard has been finished");
    GMSB_REQUIRE(dplus_all_dev != nullptr && piece_stride >= s.piece_len, "shard_finish: bad arguments");
    s.w = OrientRows();                     // the work buffers of the own range are no longer needed
    s.g->nbr.release();
    dag_from_pieces(*s.d, dplus_all_dev, pieces_dev, piece_stride, s.cut.data(), s.parts);
    Graph *g = s.g;
    g->dag = s.d;
    g->dag_pinned = true;
    s.g = nullptr;
    s.d = nullptr;
    return g;
}

Graph *graph_from_edgelist_device(int64_t m, const vid_t *src, const vid_t *dst, bool symmetrize) {
    GMSB_REQUIRE(m >= 0 && (m == 0 || (src && dst)), "graph_from_edgelist: bad arguments");
    Runtime &r = rt();
    auto *g = new Graph();
    try {
        g->directed = !symmetrize;
        int mx = 0;
        if (m) {
            DevBuf<int> d_ext(2);
            d_ext.zero();
            k_max_id<<<grid_for(m, 256), 256, 0, r.stream>>>(src, dst, m, d_ext.p); launched();
            k_min_id<<<grid_for(m, 256), 256, 0, r.stream>>>(src, dst, m, d_ext.p + 1); launched();
            mx = d_ext.get(0);
            GMSB_REQUIRE(d_ext.get(1) >= 0, "graph_from_edgelist: negative vertex id");
        }
        int64_t n = (int64_t)mx + 1;             // FindMaxNodeId(el) + 1, builder.h:285
        g->n = n;
        int64_t K = symmetrize ? 2 * m : m;
        g->off.alloc(n + 1);
        g->off.zero();
        if (K == 0) { g->slots = 0; GMSB_CUDA(cudaStreamSynchronize(r.stream)); return g; }

        DevBuf<uint64_t> keys(K), alt(K);
        k_make_keys<<<grid_for(m, 256), 256, 0, r.stream>>>(src, dst, m, symmetrize, keys.p); launched();
        uint64_t *sorted = radix_sort_keys(keys.p, alt.p, K, 0, 32 + bits_for((uint64_t)mx));

        DevBuf<uint8_t> keep(K);
        k_flag_unique<<<grid_for(K, 256), 256, 0, r.stream>>>(sorted, K, keep.p,
                                                              reinterpret_cast<unsigned long long *>(g->off.p));
        launched();
        DevBuf<int64_t> pos(K);
        exclusive_sum(keep.p, pos.p, K);
        inclusive_sum_inplace(g->off.p, n + 1);
        g->slots = g->off.get(n);
        g->nbr.alloc(g->slots);
        k_scatter_kept<<<grid_for(K, 256), 256, 0, r.stream>>>(sorted, keep.p, pos.p, K, g->nbr.p); launched();
        GMSB_CUDA(cudaStreamSynchronize(r.stream));
    } catch (...) { delete g; throw; }
    return g;
}

Graph *graph_relabel_by_degree(const Graph &in) {
    GMSB_REQUIRE(!in.directed, "relabel_by_degree: cannot relabel a directed graph");   // builder.h:1702-1705
    Runtime &r = rt();
    auto *g = new Graph();
    try {
        int64_t n = in.n, K = in.slots;
        g->n = n; g->slots = K; g->directed = false;
        g->off.alloc(n + 1);
        g->nbr.alloc(K);
        if (n == 0) return g;
        DevBuf<uint64_t> dk(n), dalt(n);
        k_degree_keys_desc<<<grid_for(n, 256), 256, 0, r.stream>>>(in.off.p, n, dk.p); launched();
        uint64_t *order = radix_sort_keys(dk.p, dalt.p, n, 0, 64, /*descending=*/true);
        DevBuf<vid_t> newid(n);
        k_newid_from_order<<<grid_for(n, 256), 256, 0, r.stream>>>(order, n, newid.p, g->off.p); launched();
        inclusive_sum_inplace(g->off.p, n + 1);
        if (K) {
            DevBuf<uint64_t> keys(K), alt(K);
            k_relabel_keys<<<grid_for(n * 32, 256), 256, 0, r.stream>>>(in.off.p, in.nbr.p, n, newid.p, keys.p);
            launched();
            uint64_t *sorted = radix_sort_keys(keys.p, alt.p, K, 0, 32 + bits_for((uint64_t)(n - 1)));
            k_low32<<<grid_for(K, 256), 256, 0, r.stream>>>(sorted, K, g->nbr.p); launched();
        }
        GMSB_CUDA(cudaStreamSynchronize(r.stream));
    } catch (...) { delete g; throw; }
    return g;
}

}  // namespace gmsb
