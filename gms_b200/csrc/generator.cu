// generator.cu — synthetic edge lists with the reference generator's exact semantics (host side, input only).
//
// Replaces  Generator::{MakeRMatEL, MakeUniformEL, PermuteIDs}   gms/third_party/gapbs/generator.h:52-114
//
// The stream is defined by libstdc++'s mt19937 / uniform_real_distribution<float> / std::shuffle, reseeded with
// kRandSeed + block every 2^18 edges, so the same standard library must produce it (SURVEY.md §7 "Bit-exact
// inputs").  Blocks are independent, so they are spread over host threads; the result is identical for any thread
// count.  This is input preparation, outside every timed region; the graph itself is then built on the GPU
// (graph_build.cu).
#include "common.cuh"

#include <algorithm>
#include <numeric>
#include <random>
#include <thread>
#include <vector>

namespace gmsb {

namespace {
constexpr int64_t kRandSeed = 27491095;            // gms/third_party/gapbs/util.h:25
constexpr int64_t kBlock = int64_t(1) << 18;       // gms/third_party/gapbs/generator.h:140

template <typename F>
void parallel_blocks(int64_t nblocks, F &&body) {
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)std::max<int64_t>(1, std::min<int64_t>(hw ? hw : 8, nblocks));
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t)
        pool.emplace_back([&, t] { for (int64_t b = t; b < nblocks; b += nt) body(b); });
    for (auto &th : pool) th.join();
}
}  // namespace

void generate_rmat(int scale, int64_t m, float A, float B, float C, bool permute, vid_t *src, vid_t *dst) {
    GMSB_REQUIRE(scale >= 1 && scale <= 30 && m >= 0 && (m == 0 || (src && dst)), "generate_rmat: bad arguments");
    const float ab = A + B, abc = A + B + C;
    const int64_t nblocks = ceil_div(m, kBlock);
    parallel_blocks(nblocks, [&](int64_t blk) {
        std::mt19937 rng;
        std::uniform_real_distribution<float> unit(0, 1.0f);
        rng.seed(kRandSeed + blk);
        const int64_t lo = blk * kBlock, hi = std::min(lo + kBlock, m);
        for (int64_t e = lo; e < hi; ++e) {
            vid_t s = 0, d = 0;
            for (int level = 0; level < scale; ++level) {
                const float p = unit(rng);
                s <<= 1; d <<= 1;
                if (p < ab) { if (p > A) d++; }
                else { s++; if (p > abc) d++; }
            }
            src[e] = s; dst[e] = d;
        }
    });
    if (permute) {
        const int64_t n = int64_t(1) << scale;
        std::vector<vid_t> perm(n);
        std::iota(perm.begin(), perm.end(), 0);
        std::mt19937 rng(kRandSeed);
        std::shuffle(perm.begin(), perm.end(), rng);
        parallel_blocks(nblocks, [&](int64_t blk) {
            const int64_t lo = blk * kBlock, hi = std::min(lo + kBlock, m);
            for (int64_t e = lo; e < hi; ++e) { src[e] = perm[src[e]]; dst[e] = perm[dst[e]]; }
        });
    }
}

void generate_uniform(int scale, int64_t m, vid_t *src, vid_t *dst) {
    GMSB_REQUIRE(scale >= 1 && scale <= 30 && m >= 0 && (m == 0 || (src && dst)), "generate_uniform: bad arguments");
    const int64_t n = int64_t(1) << scale;
    parallel_blocks(ceil_div(m, kBlock), [&](int64_t blk) {
        std::mt19937 rng;
        std::uniform_int_distribution<vid_t> pick(0, (vid_t)(n - 1));
        rng.seed(kRandSeed + blk);
        const int64_t lo = blk * kBlock, hi = std::min(lo + kBlock, m);
        for (int64_t e = lo; e < hi; ++e) {
                // Edge(udist(rng), udist(rng)): g++ evaluates constructor arguments right to left, so the FIRST
                // draw is the destination (generator.h:74)
                dst[e] = pick(rng); src[e] = pick(rng);
            }
    });
}

}  // namespace gmsb
