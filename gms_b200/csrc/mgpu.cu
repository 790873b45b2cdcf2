// mgpu.cu — several GPUs inside ONE process, below the C ABI (gmsb_set_devices; 8e).
//
// A C++ host such as the reference's benchmark harness is one process; to use the 8 GPUs of a box it calls
// gmsb_set_devices(8, ids) once and then the *_multi entry points with the handle it already has:
//   * the graph's CSR is replicated to the other devices over NVLink (cudaMemcpyPeer, cached on the handle);
//   * one host thread per device runs the partitioned form of the kernel on its replica (the same part_index /
//     part_count partition the torchrun path uses: edges by the owner of their closing vertex, clique sub-problems dealt
//     round-robin);
//   * scalar results (triangle / clique counts) are summed on the host — eight 8-byte values need no collective —
//     while array results go through NCCL: vertex_count2 is one ncclAllReduce(int64[n]) of the per-vertex partial sums,
//     the per-edge similarity one ncclAllReduce(uint32[m]) of the partial edge supports, after which every device scores
//     its own slice of the edges (edge-partitioned output, no further exchange).
// NCCL is loaded with dlopen("libnccl.so.2") the first time an array result is reduced, so the library itself has no
// link-time dependency on it.
#include "common.cuh"
#include "orient.cuh"
#include "ops.cuh"
#include "sort.cuh"

#include <cstdlib>
#include <dlfcn.h>
#include <nccl.h>

#include <exception>
#include <mutex>
#include <thread>
#include <vector>

namespace gmsb {

namespace {

struct Nccl {
    void *dll = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::vector<ncclComm_t> comms;
};

struct MultiGpu {
    std::vector<int> devs;
    Nccl nccl;
    std::mutex mutex;
};
MultiGpu g_mg;

void nccl_check(ncclResult_t r, const char *what) {
    if (r != ncclSuccess)
        throw Error(GMSB_ERR_CUDA, std::string(what) + ": " + (g_mg.nccl.GetErrorString ? g_mg.nccl.GetErrorString(r) : "NCCL error"));
}

void ensure_nccl() {
    Nccl &n = g_mg.nccl;
    if (!n.comms.empty()) return;
    if (!n.dll) {
        // GMSB_NCCL_LIB: the libnccl.so.2 to load.  A host process that also carries its own NCCL (PyTorch bundles one)
        // must point this at the same file: the loader keeps ONE object per SONAME, so whichever libnccl.so.2 is opened
        // first is the one everybody gets, and an older system copy loaded here would starve a later `import torch` of
        // symbols.  gms_b200/capi.py: set_devices() sets it to the pip-installed NCCL when there is one.
        const char *path = getenv("GMSB_NCCL_LIB");
        n.dll = dlopen(path && *path ? path : "libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!n.dll) throw Error(GMSB_ERR_UNSUPPORTED, std::string("libnccl.so.2 could not be loaded: ") + dlerror());
        auto sym = [&](const char *name) {
            void *p = dlsym(n.dll, name);
            if (!p) throw Error(GMSB_ERR_UNSUPPORTED, std::string("libnccl.so.2 lacks ") + name);
            return p;
        };
        n.CommInitAll = reinterpret_cast<decltype(n.CommInitAll)>(sym("ncclCommInitAll"));
        n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
        n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(sym("ncclAllReduce"));
        n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(sym("ncclGroupStart"));
        n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(sym("ncclGroupEnd"));
        n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
    }
    n.comms.resize(g_mg.devs.size());
    nccl_check(n.CommInitAll(n.comms.data(), (int)g_mg.devs.size(), g_mg.devs.data()), "ncclCommInitAll");
}

// f(i, device) on one host thread per device, each bound to its device; the first exception is rethrown
template <typename F>
void on_devices(F &&f) {
    const int N = (int)g_mg.devs.size();
    std::vector<std::exception_ptr> err((size_t)N);
    std::vector<std::thread> th;
    for (int i = 0; i < N; ++i)
        th.emplace_back([&, i] {
            try {
                bind_device(g_mg.devs[i]);
                f(i, g_mg.devs[i]);
            } catch (...) { err[(size_t)i] = std::current_exception(); }
        });
    for (auto &t : th) t.join();
    for (auto &e : err) if (e) std::rethrow_exception(e);
}

// in-place sum over the devices of buf[i] (count elements each); called from the host thread that owns the handle
void all_reduce(std::vector<void *> &bufs, size_t count, ncclDataType_t type) {
    ensure_nccl();
    Nccl &n = g_mg.nccl;
    nccl_check(n.GroupStart(), "ncclGroupStart");
    for (size_t i = 0; i < bufs.size(); ++i) {
        cudaSetDevice(g_mg.devs[i]);
        nccl_check(n.AllReduce(bufs[i], bufs[i], count, type, ncclSum, n.comms[i], nullptr), "ncclAllReduce");
    }
    nccl_check(n.GroupEnd(), "ncclGroupEnd");
    for (size_t i = 0; i < bufs.size(); ++i) {
        cudaSetDevice(g_mg.devs[i]);
        GMSB_CUDA(cudaStreamSynchronize(nullptr));
    }
    cudaSetDevice(current_device());
}

struct Replicas {
    std::vector<Graph *> g;          // [0] is the handle itself (not owned)
};

Replicas &replicas(Graph &g) {
    GMSB_REQUIRE(g_mg.devs.size() >= 1, "multi-GPU entry point without gmsb_set_devices");
    GMSB_REQUIRE(current_device() == g_mg.devs[0], "the graph must live on the first device of gmsb_set_devices");
    if (g.replicas) return *static_cast<Replicas *>(g.replicas);
    auto *r = new Replicas();
    r->g.assign(g_mg.devs.size(), nullptr);
    r->g[0] = &g;
    g.replicas = r;
    g.release_replicas = [](void *p) {
        auto *rr = static_cast<Replicas *>(p);
        for (size_t i = 1; i < rr->g.size(); ++i) delete rr->g[i];
        delete rr;
    };
    GMSB_CUDA(cudaStreamSynchronize(rt().stream));
    on_devices([&](int i, int dev) {
        if (i == 0) return;
        auto *c = new Graph();
        r->g[(size_t)i] = c;
        c->n = g.n; c->slots = g.slots; c->directed = g.directed;
        c->off.alloc((size_t)g.n + 1);
        c->nbr.alloc((size_t)g.slots);
        GMSB_CUDA(cudaMemcpyPeerAsync(c->off.p, dev, g.off.p, g_mg.devs[0], sizeof(eid_t) * ((size_t)g.n + 1), rt().stream));
        if (g.slots)
            GMSB_CUDA(cudaMemcpyPeerAsync(c->nbr.p, dev, g.nbr.p, g_mg.devs[0], sizeof(vid_t) * (size_t)g.slots, rt().stream));
        GMSB_CUDA(cudaStreamSynchronize(rt().stream));
    });
    return *r;
}

}  // namespace

void mg_set_devices(int n, const int *ids) {
    GMSB_REQUIRE(n >= 1 && ids != nullptr, "set_devices: bad arguments");
    int count = 0;
    GMSB_CUDA(cudaGetDeviceCount(&count));
    std::vector<int> devs(ids, ids + n);
    for (int i = 0; i < n; ++i) {
        GMSB_REQUIRE(devs[i] >= 0 && devs[i] < count, "set_devices: device index beyond the visible devices");
        for (int j = 0; j < i; ++j) GMSB_REQUIRE(devs[j] != devs[i], "set_devices: device listed twice");
    }
    std::lock_guard<std::mutex> lock(g_mg.mutex);
    for (auto c : g_mg.nccl.comms) if (c && g_mg.nccl.CommDestroy) g_mg.nccl.CommDestroy(c);
    g_mg.nccl.comms.clear();
    g_mg.devs = devs;
    for (int i = 0; i < n; ++i) {                    // peer access in both directions where the hardware allows it
        cudaSetDevice(devs[i]);
        for (int j = 0; j < n; ++j) {
            if (i == j) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devs[i], devs[j]);
            if (can) { cudaError_t e = cudaDeviceEnablePeerAccess(devs[j], 0); if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError(); }
        }
    }
    cudaSetDevice(devs[0]);
}
int mg_device_count() { return (int)g_mg.devs.size(); }

void mg_release(Graph &g) {
    if (g.replicas && g.release_replicas) g.release_replicas(g.replicas);
    g.replicas = nullptr;
}

void mg_tc_total(Graph &g, uint64_t *out) {
    GMSB_REQUIRE(out != nullptr, "tc_total: null output");
    Replicas &r = replicas(g);
    const int N = (int)r.g.size();
    std::vector<uint64_t> part((size_t)N, 0);
    on_devices([&](int i, int) {
        gmsb_tc_options opt{};
        opt.part_index = i; opt.part_count = N; opt.reuse_plan = 1;
        tc_total(*r.g[(size_t)i], opt, &part[(size_t)i], nullptr);
    });
    uint64_t total = 0;
    for (auto p : part) total += p;
    *out = total;
}

void mg_kclique_count(Graph &g, int k, uint64_t *out) {
    GMSB_REQUIRE(out != nullptr, "kclique_count: null output");
    Replicas &r = replicas(g);
    const int N = (int)r.g.size();
    std::vector<uint64_t> part((size_t)N, 0);
    on_devices([&](int i, int) { kclique_count(*r.g[(size_t)i], k, &part[(size_t)i], i, N); });
    uint64_t total = 0;
    for (auto p : part) total += p;
    *out = total;
}

// TriangleCount::Par::vertex_count2 over the devices: partial supports -> partial per-vertex sums -> one
// ncclAllReduce(int64[n]) (gms/algorithms/set_based/triangle_count/parallel/vertex.h:15-49)
void mg_tc_vertex2(Graph &g, int64_t *out_n) {
    GMSB_REQUIRE(!g.directed, "vertex_count2: graph must be undirected");
    if (g.n == 0) return;
    Replicas &r = replicas(g);
    const int N = (int)r.g.size();
    std::vector<DevBuf<unsigned long long>> t2((size_t)N);
    on_devices([&](int i, int) {
        Graph &gi = *r.g[(size_t)i];
        DevBuf<uint32_t> sup;
        tc_support(gi, sup, i, N);
        t2[(size_t)i].alloc((size_t)gi.n);
        t2[(size_t)i].zero();
        support_to_vertex2(gi, sup.p, t2[(size_t)i].p);
        GMSB_CUDA(cudaStreamSynchronize(rt().stream));
    });
    std::vector<void *> bufs;
    for (auto &b : t2) bufs.push_back(b.p);
    all_reduce(bufs, (size_t)g.n, ncclUint64);
    DevBuf<int64_t> out((size_t)g.n);
    vertex2_unrank(g, t2[0].p, out.p);
    out.download(out_n, (size_t)g.n);
    on_devices([&](int i, int) { t2[(size_t)i].release(); });          // each buffer goes back to its own device's arena
}

// One score per undirected edge u<v in CSR order, the edges partitioned over the devices by vertex ranges of equal
// edge count.  The measures that depend on the graph only through |N(a) ∩ N(b)| take the edge supports (partial per
// device, one ncclAllReduce(uint32[m])); Adamic-Adar / Resource intersect their own slice of the pairs.
void mg_edge_similarity(Graph &g, int metric, double *out, int64_t *m_out) {
    GMSB_REQUIRE(!g.directed, "edge_similarity: graph must be undirected");
    Replicas &r = replicas(g);
    const int N = (int)r.g.size();
    DevBuf<int64_t> base0;
    int64_t m = 0;
    upper_edge_base(g, base0, &m);
    if (m_out) *m_out = m;
    if (!out || m == 0) return;
    std::vector<int64_t> hbase((size_t)g.n + 1);
    base0.download(hbase.data(), (size_t)g.n + 1);
    std::vector<int64_t> cut((size_t)N + 1, g.n);
    cut[0] = 0;
    for (int i = 1; i < N; ++i)
        cut[(size_t)i] = std::lower_bound(hbase.begin(), hbase.end(), m / N * i) - hbase.begin();
    const bool from_support = metric != GMSB_SIM_ADAMIC_ADAR && metric != GMSB_SIM_RESOURCE;
    std::vector<DevBuf<uint32_t>> sup((size_t)N);
    if (from_support) {
        on_devices([&](int i, int) {
            tc_support(*r.g[(size_t)i], sup[(size_t)i], i, N);
            GMSB_CUDA(cudaStreamSynchronize(rt().stream));
        });
        std::vector<void *> bufs;
        for (auto &b : sup) bufs.push_back(b.p);
        all_reduce(bufs, sup[0].n, ncclUint32);
    }
    on_devices([&](int i, int) {
        Graph &gi = *r.g[(size_t)i];
        const int64_t a0 = cut[(size_t)i], a1 = cut[(size_t)i + 1];
        const int64_t e0 = hbase[(size_t)a0], e1 = hbase[(size_t)a1];
        if (e1 > e0) {
            DevBuf<int64_t> base;
            int64_t mi = 0;
            upper_edge_base(gi, base, &mi);
            if (from_support) {
                DevBuf<double> scores((size_t)m);
                edge_scores_range(gi, metric, base.p, sup[(size_t)i].p, a0, a1, scores.p);
                GMSB_CUDA(cudaMemcpyAsync(out + e0, scores.p + e0, sizeof(double) * (size_t)(e1 - e0), cudaMemcpyDeviceToHost,
                                          rt().stream));
                GMSB_CUDA(cudaStreamSynchronize(rt().stream));
            } else {
                DevBuf<vid_t> pa((size_t)m), pb((size_t)m);
                emit_upper_pairs(gi, base.p, pa.p, pb.p);
                pair_similarity_device(gi, metric, e1 - e0, pa.p + e0, pb.p + e0, out + e0);
            }
        }
        sup[(size_t)i].release();
    });
}

}  // namespace gmsb
