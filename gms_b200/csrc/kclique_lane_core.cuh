// kclique_lane_core.cuh — the per-lane part of the k-clique search (host/device, no warp primitives).
//
// Replaces the inner recursion of KClique::KcListing::listing (gms/algorithms/non_set_based/k_clique_list/kernels/
// kclisting.h:92-114): there a recursion level relabels and swaps adjacency entries and the last level counts one
// clique at a time; here a sub-problem is a COMPACT bit matrix (c <= 512 members, row a = members after a that are
// adjacent to a) and every GPU lane runs its own depth-first search over it with the candidate sets held in
// registers.  A level is one AND of NW 64-bit words, the last two levels are one pass of AND + popcount.
//
// The functions are __host__ __device__ so that tests/cpp/lane_core_test.cpp can run exactly this code on the CPU
// against a brute-force count (the container that builds the library has no GPU).
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define GMSB_HD __host__ __device__ __forceinline__
#else
#define GMSB_HD inline
#endif

namespace gmsb {
namespace lane {

using u64 = unsigned long long;

constexpr int kCMax = 512;         // members of a compact sub-problem (9-bit member indices)
constexpr int kPathBits = 9;
constexpr int kPathLevels = 7;     // 7 x 9 bits in one 64-bit register
constexpr int kMaxNeed = kPathLevels + 2;   // deepest stored level is need - 3
constexpr int kC3Max = 192;        // members of a per-warp third-level matrix (kclique_lane.cuh: warp_tasks)
constexpr int kP3 = 3;             // its pitch: up to 3 valid 64-bit words per row

GMSB_HD int popc64(u64 x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
GMSB_HD int ctz64(u64 x) {     // x != 0
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}

// pitch (64-bit words per row) of a compact matrix with NW valid words per row: odd, so that lanes reading the same
// word of 32 different rows spread over the shared-memory banks
GMSB_HD constexpr int pitch_for(int nw) { return nw | 1; }

// bits b of a 64-bit word with (b mod 2^split_log2) == s: the second clique vertex of a task is restricted to one
// residue class, which cuts a member's subtree into 2^split_log2 independent tasks without any index arithmetic
GMSB_HD u64 stripe_mask(int split_log2, int s) {
    u64 rep;
    switch (split_log2) {
        case 0: rep = ~0ull; break;
        case 1: rep = 0x5555555555555555ull; break;
        case 2: rep = 0x1111111111111111ull; break;
        case 3: rep = 0x0101010101010101ull; break;
        case 4: rep = 0x0001000100010001ull; break;
        case 5: rep = 0x0000000100000001ull; break;
        default: rep = 1ull; break;
    }
    return rep << s;
}

// bits of word w (bit positions 64w .. 64w+63) whose position is greater than `pos`
GMSB_HD u64 above_mask(int w, int pos) {
    const int rel = pos - (w << 6);
    if (rel < 0) return ~0ull;
    if (rel >= 63) return 0ull;
    return ~0ull << (rel + 1);
}

template <int NW>
struct LaneState {
    u64 cur[NW];      // candidate set at the current level
    u64 path;         // kPathBits-bit fields: [0] = first member, [l] = member picked at level l-1
    u64 stripe;       // residue-class mask applied to the picks of level 0
    int pos;          // last member tried at the current level (members are tried in ascending order)
    int level;        // -1 = idle
};

// task = (first member a, residue class `cls` of the second member)
template <int NW>
GMSB_HD void lane_begin(LaneState<NW> &s, const u64 *cm, int pitch, int a, int cls, int split_log2) {
    s.stripe = stripe_mask(split_log2, cls);
    const u64 *row = cm + (size_t)a * pitch;
#pragma unroll
    for (int w = 0; w < NW; ++w) s.cur[w] = row[w];
    s.level = 0;
    s.pos = a;                 // row a only has members after a
    s.path = (u64)a;
}

// Walks the search tree of the lane's task until it reaches a candidate set Q with exactly two vertices left to
// pick (returns true; the caller counts the pairs inside Q) or the task is finished (returns false, level = -1).
// need >= 4: size of the cliques counted inside the compact graph (the first member is one of them).
// Per level only the picked member is remembered (in `path`); the parent's candidate set is recomputed from the
// rows on the path on the way back, so the search needs no per-lane stack memory.
template <int NW>
GMSB_HD bool lane_advance(LaneState<NW> &s, const u64 *cm, int pitch, int need, u64 (&Q)[NW]) {
    for (;;) {
        // next member of cur after pos (level 0: inside the task's residue class)
        const u64 sm = s.level == 0 ? s.stripe : ~0ull;
        int v = -1;
#pragma unroll
        for (int w = NW - 1; w >= 0; --w) {
            const u64 m = s.cur[w] & sm & above_mask(w, s.pos);
            if (m) v = (w << 6) + ctz64(m);
        }
        if (v < 0) {
            if (s.level == 0) { s.level = -1; return false; }
            s.pos = (int)((s.path >> (kPathBits * s.level)) & (kCMax - 1));     // resume after the child just left
            --s.level;
            const u64 *r0 = cm + (size_t)(s.path & (kCMax - 1)) * pitch;
#pragma unroll
            for (int w = 0; w < NW; ++w) s.cur[w] = r0[w];
            for (int t = 1; t <= s.level; ++t) {
                const u64 *r = cm + (size_t)((s.path >> (kPathBits * t)) & (kCMax - 1)) * pitch;
#pragma unroll
                for (int w = 0; w < NW; ++w) s.cur[w] &= r[w];
            }
            continue;
        }
        s.pos = v;
        const u64 *row = cm + (size_t)v * pitch;
        int pc = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            Q[w] = s.cur[w] & row[w];
            pc += popc64(Q[w]);
        }
        const int left = need - s.level - 2;         // vertices still to pick after v
        if (left == 2) {
            if (pc >= 2) return true;
        } else if (pc >= left) {
            ++s.level;
            const int sh = kPathBits * s.level;
            s.path = (s.path & ~((u64)(kCMax - 1) << sh)) | ((u64)v << sh);
#pragma unroll
            for (int w = 0; w < NW; ++w) s.cur[w] = Q[w];
        }
    }
}

// sum over x in I of |A ∩ row(x)|: the last two clique vertices in one pass (row(x) only has members after x, so
// the words below x's own are skipped).  A variant that added the rows of seven members as bit planes (carry-save
// adders, 3 popcounts per word instead of 7) was measured on B200 and was SLOWER (scale-22 k=6: 44 s vs 38 s): the
// candidate sets this deep hold only 3-5 members per 64-bit word, so the groups were mostly padding.  So was a
// single loop over all members with the words below x's own predicated off (68 s vs 36 s): a predicated-off POPC
// still takes its slot on the quarter-rate pipe.
template <int NW>
GMSB_HD unsigned leaf_pairs(const u64 *cm, int pitch, const u64 (&I)[NW], const u64 (&A)[NW]) {
    unsigned cnt = 0;
#pragma unroll
    for (int w1 = 0; w1 < NW; ++w1) {
        u64 bits = I[w1];
        while (bits) {
            const int x = (w1 << 6) + ctz64(bits);
            bits &= bits - 1;
            const u64 *row = cm + (size_t)x * pitch;
#pragma unroll
            for (int w = w1; w < NW; ++w) cnt += (unsigned)popc64(A[w] & row[w]);
        }
    }
    return cnt;
}

// whole task on one lane (the CPU test and the GPU lanes run the same sequence of calls)
template <int NW>
GMSB_HD u64 lane_run_task(const u64 *cm, int pitch, int need, unsigned task, int split_log2) {
    LaneState<NW> s;
    lane_begin<NW>(s, cm, pitch, (int)(task >> split_log2), (int)(task & ((1u << split_log2) - 1u)), split_log2);
    u64 Q[NW];
    if (need == 3) {           // the task is one pass: second member in the residue class, third anywhere in the row
#pragma unroll
        for (int w = 0; w < NW; ++w) Q[w] = s.cur[w] & s.stripe;
        return leaf_pairs<NW>(cm, pitch, Q, s.cur);
    }
    u64 total = 0;
    while (lane_advance<NW>(s, cm, pitch, need, Q)) total += leaf_pairs<NW>(cm, pitch, Q, Q);
    return total;
}

// ---- flat form of the lane search ----------------------------------------------------------------------------------
// lane_run_task alternates between the tree walk (lane_advance) and the leaf pass (leaf_pairs), and the lanes of a
// warp are at different points of that alternation almost all the time: ncu showed 12 of 32 lanes active per
// instruction (round 1).  The flat form makes ONE leaf operation the unit of every loop iteration:
//     job  = (A, I): count sum over x in I of |A ∩ row(x)|        (I = A, or A ∩ residue class for a need == 3 task)
//     step = take the next x of I, AND its row with A, popcount    (identical instruction sequence on every lane)
// and moves everything else — walking the tree to the next job, pulling the next task — into flat_refill, which the
// warp runs only when enough lanes have run dry (kclique_lane.cuh: lane_tasks_flat), so that its cost is shared.
// All NW words of a row are visited (rows only have members after x, the lower words AND to zero): the flat form is
// used on compact matrices of <= 4 words per row, where skipping them saves less than the divergence costs.
template <int NW>
struct FlatState {
    LaneState<NW> dfs;    // the walk above the jobs (need >= 4); level -1 = no task in progress
    u64 A[NW];            // the job's set
    u64 imask;            // I = A & imask
    u64 bits;             // members of I still to do in word w1
    int w1;               // current word of I; NW = no job
    bool exhausted;       // the task ticket has run out for this lane
};

template <int NW>
GMSB_HD void flat_init(FlatState<NW> &s) {
    s.dfs.level = -1;
    s.bits = 0;
    s.w1 = NW;
    s.exhausted = false;
}
// moves to the next non-empty word of I; afterwards the job has a member to do iff s.w1 < NW
template <int NW>
GMSB_HD void flat_skip(FlatState<NW> &s) {
    while (s.bits == 0 && s.w1 < NW) {
        ++s.w1;
#pragma unroll
        for (int w = 1; w < NW; ++w)
            if (w == s.w1) s.bits = s.A[w] & s.imask;
    }
}
// one leaf operation (the caller checked s.w1 < NW after flat_skip)
template <int NW>
GMSB_HD unsigned flat_step(FlatState<NW> &s, const u64 *cm, int pitch) {
    const int x = (s.w1 << 6) + ctz64(s.bits);
    s.bits &= s.bits - 1;
    const u64 *row = cm + (size_t)x * pitch;
    unsigned cnt = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) cnt += (unsigned)popc64(s.A[w] & row[w]);
    return cnt;
}
GMSB_HD unsigned flat_fetch(unsigned *counter) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(counter, 1u);
#else
    return (*counter)++;
#endif
}
// Gives a dry lane its next job: continues the lane's tree walk, or pulls tasks [0, t_end) from *counter until one
// yields a job.  Task t = (first member t >> split_log2 — alist[t >> split_log2] when a list is given —, residue
// class of the second member).  Sets s.exhausted when there is nothing left.
template <int NW>
GMSB_HD void flat_refill(FlatState<NW> &s, const u64 *cm, int pitch, int need, int split_log2, unsigned t_end,
                         unsigned *counter, const unsigned short *alist) {
    for (;;) {
        if (s.dfs.level >= 0 && lane_advance<NW>(s.dfs, cm, pitch, need, s.A)) {
            s.imask = ~0ull;
            s.w1 = 0;
            s.bits = s.A[0];
            return;
        }
        const unsigned t = flat_fetch(counter);
        if (t >= t_end) { s.exhausted = true; return; }
        const unsigned m = t >> split_log2;
        lane_begin<NW>(s.dfs, cm, pitch, alist ? (int)alist[m] : (int)m, (int)(t & ((1u << split_log2) - 1u)),
                       split_log2);
        if (need == 3) {               // the task is one job: second member in the residue class, third anywhere
#pragma unroll
            for (int w = 0; w < NW; ++w) s.A[w] = s.dfs.cur[w];
            s.imask = s.dfs.stripe;
            s.dfs.level = -1;
            s.w1 = 0;
            s.bits = s.A[0] & s.imask;
            return;
        }
    }
}
// refill policy: the dry lanes are served when they are a quarter of the lanes that still have work, or nobody has a job
GMSB_HD bool flat_should_refill(int dry_lanes, int live_lanes) {
    return dry_lanes > 0 && (live_lanes == 0 || 4 * dry_lanes >= dry_lanes + live_lanes);
}

// ---- compaction of a candidate set of a big matrix into a compact matrix -------------------------------------------
// set: P1 words over the positions of the big matrix; prefix[w] = members in words < w
GMSB_HD int compact_index(const u64 *set, const int *prefix, int p) {
    const int w = p >> 6;
    return prefix[w] + popc64(set[w] & ((1ull << (p & 63)) - 1ull));
}

// compact row of the member at position pa (big row `row`, W1 valid words): bit idx(q) for every member q in the row
GMSB_HD void compact_row(const u64 *set, const int *prefix, int W1, const u64 *row, int pa, u64 *out, int nwb) {
    for (int w2 = 0; w2 < nwb; ++w2) out[w2] = 0ull;
    int curw = -1;
    u64 acc = 0;
    for (int w = pa >> 6; w < W1; ++w) {
        const u64 sw = set[w];
        u64 x = row[w] & sw;
        const int base = prefix[w];
        while (x) {
            const int b = ctz64(x);
            x &= x - 1;
            const int idx = base + popc64(sw & ((1ull << b) - 1ull));
            if ((idx >> 6) != curw) {
                if (curw >= 0) out[curw] = acc;
                curw = idx >> 6;
                acc = 0;
            }
            acc |= 1ull << (idx & 63);
        }
    }
    if (curw >= 0) out[curw] = acc;
}

}  // namespace lane
}  // namespace gmsb
