// common.cuh — runtime plumbing shared by every translation unit of libgmsb.so:
// error handling, the per-process stream, launch accounting and an RAII device buffer.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <utility>

#include "../../include/gmsb.h"

namespace gmsb {

using vid_t = int32_t;     // vertex id (gms/common/types.h:9)
using eid_t = int64_t;     // CSR offset

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define GMSB_CUDA(expr)                                                                                  \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            throw ::gmsb::Error(e_ == cudaErrorMemoryAllocation ? GMSB_ERR_OOM : GMSB_ERR_CUDA,          \
                                std::string(#expr) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ +   \
                                    ":" + std::to_string(__LINE__) + ")");                               \
        }                                                                                                \
    } while (0)

#define GMSB_REQUIRE(cond, msg)                                        \
    do {                                                               \
        if (!(cond)) throw ::gmsb::Error(GMSB_ERR_INVALID, (msg));     \
    } while (0)

struct Runtime {
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    int sm_count = 0;
    int device = -1;
    size_t smem_optin = 0;
};
Runtime &rt();             // of the calling thread's device; lazily initialised; throws GMSB_ERR_CUDA without a device
void set_last_error(const std::string &msg);
const std::string &last_error_message();
int current_device();      // the calling thread's bound device, else the process's primary device
void bind_device(int device);      // bind the calling host thread to a device (-1: back to the primary one)
void *arena_alloc(size_t bytes);   // caching device allocator (size-class free lists over cudaMalloc)
void arena_free(void *p);
void arena_trim();                 // give every cached block back to the driver

// Count a kernel launch (bench.py reports the total as gpu_launches) and surface launch errors early.
inline void launched() {
    rt().launches++;
    GMSB_CUDA(cudaPeekAtLastError());
}

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    // Allocation goes through the library's caching arena (capi.cu: arena_alloc / arena_free): freed blocks are
    // kept in size-class free lists and handed out again, so the multi-GB scratch of a graph build is recycled by
    // the next build instead of being unmapped and re-mapped by the driver — cudaMalloc/cudaFree (and the
    // re-mapping a cudaMallocAsync pool does when sizes change) cost more than the kernels they feed.  All work is
    // on one stream, so reuse of a freed block is ordered after its last use.
    void alloc(size_t count) {
        release();
        n = count;
        if (count) p = static_cast<T *>(arena_alloc(count * sizeof(T)));
    }
    void release() {
        if (p) arena_free(p);
        p = nullptr; n = 0;
    }
    void zero() { if (n) GMSB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), rt().stream)); }
    void upload(const T *host, size_t count) {
        if (count) GMSB_CUDA(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, rt().stream));
    }
    void download(T *host, size_t count) const {
        if (count) GMSB_CUDA(cudaMemcpyAsync(host, p, count * sizeof(T), cudaMemcpyDeviceToHost, rt().stream));
        GMSB_CUDA(cudaStreamSynchronize(rt().stream));
    }
    T get(size_t i) const {
        T v;
        GMSB_CUDA(cudaMemcpyAsync(&v, p + i, sizeof(T), cudaMemcpyDeviceToHost, rt().stream));
        GMSB_CUDA(cudaStreamSynchronize(rt().stream));
        return v;
    }
};

// Device timer on the library's stream.
struct DevTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    DevTimer() { cudaEventCreate(&a); cudaEventCreate(&b); }
    ~DevTimer() { cudaEventDestroy(a); cudaEventDestroy(b); }
    void start() { cudaEventRecord(a, rt().stream); }
    void stop() { cudaEventRecord(b, rt().stream); }
    double ms() {
        cudaEventSynchronize(b);
        float t = 0;
        cudaEventElapsedTime(&t, a, b);
        return t;
    }
};

// GMSB_TC_TRACE=1: device time of every phase of the representation build on stderr (profiling aid)
struct PhaseTrace {
    bool on;
    cudaEvent_t last = nullptr;
    explicit PhaseTrace(const char *env) : on(std::getenv(env) != nullptr) {
        if (on) { cudaEventCreate(&last); cudaEventRecord(last, rt().stream); }
    }
    ~PhaseTrace() { if (last) cudaEventDestroy(last); }
    void mark(const char *what) {
        if (!on) return;
        cudaEvent_t now;
        cudaEventCreate(&now);
        cudaEventRecord(now, rt().stream);
        cudaEventSynchronize(now);
        float ms = 0;
        cudaEventElapsedTime(&ms, last, now);
        std::fprintf(stderr, "[gmsb trace] %-28s %8.3f ms\n", what, ms);
        cudaEventDestroy(last);
        last = now;
    }
};

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int bits_for(uint64_t x) { int b = 0; while (b < 64 && (x >> b)) ++b; return b ? b : 1; }

// ---- device graph ----------------------------------------------------------------------------------------------
struct TcPlan;      // tc.cu
struct Dag;         // orient.cu

struct Graph {
    int64_t n = 0;
    int64_t slots = 0;          // CSR entries
    bool directed = false;
    DevBuf<eid_t> off;          // n+1
    DevBuf<vid_t> nbr;          // slots, ascending within each list
    Dag *dag = nullptr;         // cached degree-oriented DAG (undirected graphs only)
    bool dag_pinned = false;    // the DAG was built with the graph (GMSB_BUILD_ORIENT): reuse_plan = 0 keeps it
    bool dag_only = false;      // sharded build: only the oriented representation is complete on this device (nbr is not)
    void *replicas = nullptr;   // copies on the other devices of gmsb_set_devices (mgpu.cu)
    void (*release_replicas)(void *) = nullptr;
    ~Graph();
};

}  // namespace gmsb
