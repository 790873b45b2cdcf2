// kclique_lane.cuh — lane-parallel k-clique kernels (included by kclique.cu).
//
// Replaces: KClique::KcListing::listing + Parallelize::{node,edge}
//           (gms/algorithms/non_set_based/k_clique_list/kernels/kclisting.h:92-114,
//            parallelizationStrategy/parallelize.h:39-121).
//
// The warp-cooperative kernels in kclique.cu spend a whole warp on every search-tree node; deep in the tree the
// candidate sets have a few dozen members and most lanes idle.  Here every LANE owns a subtree:
//   * the sub-problem of a vertex u is its |S| x |S| bit matrix (S = N+(u)) in shared memory, rows of 64-bit words
//     with an odd pitch (conflict-free when 32 lanes read the same word of 32 different rows);
//   * 33 <= d+(u) <= 512 : the matrix itself is the compact problem; tasks (second vertex, residue class of the
//     third vertex) are dealt to the lanes of the CTA through a shared ticket;
//   * d+(u) > 512 : the CTA walks the top of the tree together (one AND per level, a handful of barriers) until the
//     candidate set has <= 512 members, re-indexes that set into a second, compact matrix (<= 8 words per row) and
//     deals its tasks to the lanes; with two vertices left the pairs are counted straight off the big matrix;
//   * 129 <= d+(u) <= 512 and four or more vertices left below the second one: a warp re-indexes the second
//     vertex's row once more into its own small matrix (warp_tasks).
// A lane keeps its candidate sets in registers (kclique_lane_core.cuh); only the path is remembered per level, the
// parent's set is recomputed on the way back, so there is no per-lane stack memory at all.
#pragma once
#include "common.cuh"
#include "isect.cuh"
#include "kclique_lane_core.cuh"

namespace gmsb {
namespace lane {

__device__ __forceinline__ int split_for(int c, int block) {
    int sl = 0;
    while (sl < 6 && (c << sl) < 16 * block) ++sl;       // >= 16 tasks per lane keeps the tail short
    return sl;
}

// Lanes of the calling warp(s) pull tasks [0, t_end) of the compact graph cm from *counter (zeroed by the caller) and
// count the `need`-cliques (need >= 3) of those tasks.  Task t = (first member, residue class of the second member,
// kclique_lane_core.cuh): the member is t >> split_log2, or alist[t >> split_log2] when a list of members is given.
// All 32 lanes of every calling warp must be here.  Returns this lane's partial count.
template <int NW>
__device__ __forceinline__ u64 lane_tasks(const u64 *cm, int pitch, unsigned t_end, int need, int split_log2,
                                          unsigned *counter, int lane, const unsigned short *alist) {
    u64 total = 0;
    LaneState<NW> s;
    s.level = -1;
    bool exhausted = false;
    for (;;) {
        const bool want = s.level < 0 && !exhausted;
        const unsigned idle = __ballot_sync(0xffffffffu, want);
        if (idle) {
            const int leader = __ffs(idle) - 1;
            unsigned base = 0;
            if (lane == leader) base = atomicAdd(counter, (unsigned)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (want) {
                const unsigned t = base + (unsigned)__popc(idle & ((1u << lane) - 1u));
                if (t >= t_end) exhausted = true;
                else {
                    const unsigned m = t >> split_log2;
                    lane_begin<NW>(s, cm, pitch, alist ? (int)alist[m] : (int)m, (int)(t & ((1u << split_log2) - 1u)),
                                   split_log2);
                }
            }
        }
        if (__all_sync(0xffffffffu, s.level < 0)) break;
        if (s.level >= 0) {
            u64 Q[NW];
            if (need == 3) {
#pragma unroll
                for (int w = 0; w < NW; ++w) Q[w] = s.cur[w] & s.stripe;
                total += leaf_pairs<NW>(cm, pitch, Q, s.cur);
                s.level = -1;
            } else {
                if (lane_advance<NW>(s, cm, pitch, need, Q)) total += leaf_pairs<NW>(cm, pitch, Q, Q);
            }
        }
    }
    return total;
}

// ---- third level: per-warp compact matrix ------------------------------------------------------------------------------
// With four or more vertices still to pick below a member a of cm, the search below a runs in a matrix of its own:
// the WARP re-indexes row a (<= kC3Max = 192 members) into M3 (rows of 1-3 words instead of up to 8) and its lanes pull M3's
// tasks from a warp-local ticket.  Deep in the tree the sets hold 3-5 members per word of the parent's index space, so
// every AND + popcount there is mostly zeros; one more re-indexing halves to quarters the words per step and fills
// them.  Rows that are too large for M3 are put on a list and searched in cm afterwards by all lanes of the CTA.
// Used by the 129..512 classes (scale 22, k = 6: the <= 512 class 4.53 -> 3.43 s; scale 20, k = 7: 66.9 -> 51.5 s when M3 grew from 128 to 192 members).  In the
// d+ > 512 kernel it was NOT a win (k = 7: +18 %): rows of M2 mostly exceed 128 members, a warp's 32 lanes drain at
// the end of every member, and the extra live state pushed that kernel into register spills — it deals every task of
// M2 to the lanes of the whole CTA instead.
struct WarpBox {
    u64 m3[kC3Max * kP3];
    u64 set[8];
    int prefix[8];
    unsigned short list[kC3Max];
    unsigned counter;
    unsigned pad;
};
struct CtaBigRows {                          // members of cm whose row does not fit M3
    unsigned short list[kCMax];
    unsigned count;
    unsigned counter;                        // ticket of the lane tasks over these rows
};

// nw = valid 64-bit words per row of cm (<= 8); every warp of the CTA calls this with the same arguments
__device__ u64 warp_tasks(const u64 *cm, int pitch, int nw, int c, int need, unsigned *counter, WarpBox *box,
                          CtaBigRows *big, int lane) {
    u64 total = 0;
    for (;;) {
        unsigned a = 0;
        if (lane == 0) a = atomicAdd(counter, 1u);
        a = __shfl_sync(0xffffffffu, a, 0);
        if (a >= (unsigned)c) break;
        const u64 *row = cm + (size_t)a * pitch;
        const u64 mine = lane < nw ? row[lane] : 0ull;
        const int pc = __popcll(mine);
        int incl = pc;
        for (int o = 1; o < 8; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += x;
        }
        const int c3 = __shfl_sync(0xffffffffu, incl, 7);         // lanes >= nw add 0
        if (c3 < need - 1) continue;
        if (c3 > kC3Max) {
            if (lane == 0) big->list[atomicAdd(&big->count, 1u)] = (unsigned short)a;
            continue;
        }
        __syncwarp();
        if (lane == 0) box->counter = 0u;
        if (lane < 8) { box->set[lane] = mine; box->prefix[lane] = incl - pc; }
        __syncwarp();
        for (int p = lane; p < nw * 64; p += 32)
            if ((box->set[p >> 6] >> (p & 63)) & 1ull)
                box->list[compact_index(box->set, box->prefix, p)] = (unsigned short)p;
        __syncwarp();
        const int nw3 = (c3 + 63) >> 6;              // 1..3 words per row of M3
        for (int m = lane; m < c3; m += 32) {
            const int pa = box->list[m];
            compact_row(box->set, box->prefix, nw, cm + (size_t)pa * pitch, pa, box->m3 + (size_t)m * kP3, nw3);
        }
        __syncwarp();
        const int sl = split_for(c3, 32);
        const unsigned t3 = (unsigned)c3 << sl;
        total += nw3 == 1   ? lane_tasks<1>(box->m3, kP3, t3, need - 1, sl, &box->counter, lane, nullptr)
                 : nw3 == 2 ? lane_tasks<2>(box->m3, kP3, t3, need - 1, sl, &box->counter, lane, nullptr)
                            : lane_tasks<3>(box->m3, kP3, t3, need - 1, sl, &box->counter, lane, nullptr);
        __syncwarp();
    }
    return total;
}

// `need`-cliques (need >= 3) of the compact graph cm[c][pitch] (NW valid words per row) by ALL threads of the CTA;
// *counter and *big are zeroed by the caller (and a barrier passed).  Per-lane tasks over every member, or — when the
// search below a member is deep enough to pay for another re-indexing — per-warp third-level matrices first and
// per-lane tasks over the rows that did not fit one.
template <int NW>
__device__ __forceinline__ u64 count_compact(const u64 *cm, int pitch, int c, int need, int block, unsigned *counter,
                                             WarpBox *box, CtaBigRows *big, int lane) {
    u64 total = 0;
    unsigned t_end = (unsigned)c;
    int sl = split_for(c, block);
    const unsigned short *alist = nullptr;
    if (need >= 5 && box) {
        total = warp_tasks(cm, pitch, NW, c, need, counter, box, big, lane);
        __syncthreads();
        t_end = big->count;
        sl = 6;
        alist = big->list;
        counter = &big->counter;
    }
    return total + lane_tasks<NW>(cm, pitch, t_end << sl, need, sl, counter, lane, alist);
}

// rows of the bit matrix of S (|S| = D): warp per member i streams N+(S[i]) and looks every element up in S
template <int BLOCK>
__device__ __forceinline__ void build_rows(const vid_t *S, int D, u64 *M, int pitch, const eid_t *__restrict__ off,
                                           const vid_t *__restrict__ nbr, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    uint32_t *M32 = reinterpret_cast<uint32_t *>(M);
    for (int i = warp; i < D; i += BLOCK / 32) {
        const vid_t vi = S[i];
        const eid_t mb = off[vi];
        const int md = (int)(off[vi + 1] - mb);
        for (int j = lane; j < md; j += 32) {
            const vid_t w = nbr[mb + j];
            int lo = i + 1, hi = D;                     // members after i only (ids ascend with position)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (S[mid] < w) lo = mid + 1; else hi = mid;
            }
            if (lo < D && S[lo] == w) atomicOr(&M32[(size_t)i * pitch * 2 + (lo >> 5)], 1u << (lo & 31));
        }
    }
}

// ---- 33 <= d+ <= 64*NWB ---------------------------------------------------------------------------------------------
template <int NWB, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
k_kclique_lane_mid(const vid_t *__restrict__ verts, int64_t count, const eid_t *__restrict__ off,
                   const vid_t *__restrict__ nbr, int k, unsigned long long *__restrict__ total,
                   unsigned int *__restrict__ ticket, int pi, int P, int flags) {
    constexpr int DMAX = 64 * NWB;
    constexpr int PITCH = pitch_for(NWB);
    __shared__ vid_t S[DMAX];
    __shared__ u64 M[DMAX * PITCH];
    __shared__ unsigned long long red[BLOCK / 32];
    __shared__ unsigned int s_item, s_counter;
    __shared__ CtaBigRows s_big;
    extern __shared__ u64 dyn64[];                  // one WarpBox per warp when the third level is enabled (flags & 1)
    const int tid = threadIdx.x, lane = tid & 31;
    WarpBox *box = (flags & 1) ? reinterpret_cast<WarpBox *>(dyn64) + (tid >> 5) : nullptr;
    u64 acc = 0;
    for (;;) {
        __syncthreads();
        if (tid == 0) { s_item = atomicAdd(ticket, 1u); s_counter = 0; s_big.count = 0; s_big.counter = 0; }
        __syncthreads();
        const int64_t t = pi + (int64_t)s_item * P;
        if (t >= count) break;
        const vid_t u = verts[t];
        const eid_t ob = off[u];
        const int D = (int)(off[u + 1] - ob);
        for (int j = tid; j < D; j += BLOCK) S[j] = nbr[ob + j];
        for (int j = tid; j < D * PITCH; j += BLOCK) M[j] = 0ull;
        __syncthreads();
        build_rows<BLOCK>(S, D, M, PITCH, off, nbr, tid);
        __syncthreads();
        // u is the first clique vertex; k-1 more inside the matrix
        acc += count_compact<NWB>(M, PITCH, D, k - 1, BLOCK, &s_counter, box, &s_big, lane);
    }
    const unsigned long long s = block_sum(acc, red);
    if (tid == 0 && s) atomicAdd(total, s);
}

// ---- d+ > 512 ---------------------------------------------------------------------------------------------------------
constexpr int kStackLevels = 16;

__host__ __device__ inline int huge_pitch(int maxD) { return ((maxD + 63) >> 6) | 1; }
// dynamic shared memory of k_kclique_lane_huge in 64-bit words (matrix included unless it is spilled to global)
__host__ __device__ inline size_t huge_smem_words(int maxD, bool matrix_in_smem) {
    const size_t P1 = (size_t)huge_pitch(maxD);
    return (matrix_in_smem ? (size_t)maxD * P1 : 0) + (size_t)kCMax * pitch_for(8) + (size_t)kStackLevels * P1 +
           (size_t)((maxD + 1) >> 1) + ((P1 + 2) >> 1) + (size_t)(kCMax / 4);
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, 1)
k_kclique_lane_huge(const vid_t *__restrict__ verts, const int64_t *__restrict__ item_base, int64_t nverts,
                    int64_t count, const eid_t *__restrict__ off, const vid_t *__restrict__ nbr, int k, int maxD,
                    unsigned long long *__restrict__ total, unsigned int *__restrict__ ticket,
                    u64 *__restrict__ spill, int pi, int P) {
    extern __shared__ u64 smem64[];
    const int P1 = huge_pitch(maxD);
    u64 *M1 = spill ? spill + (size_t)blockIdx.x * ((size_t)maxD * P1) : smem64;
    u64 *sp = smem64 + (spill ? 0 : (size_t)maxD * P1);
    u64 *M2 = sp;                  sp += (size_t)kCMax * pitch_for(8);
    u64 *stack = sp;               sp += (size_t)kStackLevels * P1;
    vid_t *S = reinterpret_cast<vid_t *>(sp);            sp += (maxD + 1) >> 1;
    int *prefix = reinterpret_cast<int *>(sp);           sp += (P1 + 2) >> 1;
    unsigned short *list = reinterpret_cast<unsigned short *>(sp);
    __shared__ unsigned long long red[BLOCK / 32];
    __shared__ unsigned int s_item, s_counter;
    __shared__ int s_c;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    u64 acc = 0;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(ticket, 1u);
        __syncthreads();
        const int64_t item = pi + (int64_t)s_item * P;
        if (item >= count) break;
        int64_t lo_t = 0, hi_t = nverts;
        while (hi_t - lo_t > 1) {
            const int64_t mid = (lo_t + hi_t) >> 1;
            if (item_base[mid] <= item) lo_t = mid; else hi_t = mid;
        }
        const int64_t t = lo_t;
        const int part = (int)(item - item_base[t]), nparts = (int)(item_base[t + 1] - item_base[t]);
        const vid_t u = verts[t];
        const eid_t ob = off[u];
        const int D = (int)(off[u + 1] - ob);
        const int W1 = (D + 63) >> 6;
        for (int j = tid; j < D; j += BLOCK) S[j] = nbr[ob + j];
        for (size_t j = tid; j < (size_t)D * P1; j += BLOCK) M1[j] = 0ull;
        __syncthreads();
        build_rows<BLOCK>(S, D, M1, P1, off, nbr, tid);
        // the CTA walks the top of the tree together: stack[level] = candidates after (level + 2) chosen vertices
        for (int i = part; i < D; i += nparts) {
            __syncthreads();
            for (int w = tid; w < P1; w += BLOCK) stack[w] = M1[(size_t)i * P1 + w];
            __syncthreads();
            int level = 0;
            bool fresh = true;
            int cursor[kStackLevels];
            for (;;) {
                const u64 *set = stack + (size_t)level * P1;
                if (fresh) {
                    fresh = false;
                    if (warp == 0) {
                        int pc = 0;
                        for (int w = lane; w < W1; w += 32) pc += __popcll(set[w]);
                        for (int o = 16; o; o >>= 1) pc += __shfl_xor_sync(0xffffffffu, pc, o);
                        if (lane == 0) s_c = pc;
                    }
                    __syncthreads();
                    const int c = s_c;
                    const int need = k - level - 2;
                    bool expand = false;
                    if (c >= need) {
                        if (need == 1) {
                            if (tid == 0) acc += (u64)c;
                        } else if (need == 2) {
                            // pairs straight off the big matrix: thread per member
                            unsigned cnt = 0;
                            for (int p = tid; p < D; p += BLOCK) {
                                const int w0 = p >> 6;
                                if ((set[w0] >> (p & 63)) & 1ull) {
                                    const u64 *row = M1 + (size_t)p * P1;
                                    for (int w = w0; w < W1; ++w) cnt += (unsigned)__popcll(set[w] & row[w]);
                                }
                            }
                            acc += cnt;
                        } else if (c <= kCMax) {
                            // re-index the set into the compact matrix and deal its subtrees to the lanes
                            if (warp == 0) {
                                int carry = 0;
                                for (int base = 0; base < P1; base += 32) {
                                    const int w = base + lane;
                                    const int v = w < W1 ? __popcll(set[w]) : 0;
                                    int incl = v;
                                    for (int o = 1; o < 32; o <<= 1) {
                                        const int x = __shfl_up_sync(0xffffffffu, incl, o);
                                        if (lane >= o) incl += x;
                                    }
                                    if (w < P1) prefix[w] = carry + incl - v;
                                    carry += __shfl_sync(0xffffffffu, incl, 31);
                                }
                            }
                            if (tid == 0) s_counter = 0;
                            __syncthreads();
                            for (int p = tid; p < D; p += BLOCK)
                                if ((set[p >> 6] >> (p & 63)) & 1ull)
                                    list[compact_index(set, prefix, p)] = (unsigned short)p;
                            __syncthreads();
                            const int nwb = (c + 63) >> 6;         // 1..8 words per compact row
                            const int pitch2 = pitch_for(nwb);
                            for (int a = tid; a < c; a += BLOCK) {
                                const int pa = list[a];
                                compact_row(set, prefix, W1, M1 + (size_t)pa * P1, pa, M2 + (size_t)a * pitch2, nwb);
                            }
                            __syncthreads();
                            const int sl = split_for(c, BLOCK);
                            switch (nwb) {
                                case 1: acc += lane_tasks<1>(M2, pitch2, (unsigned)c << sl, need, sl, &s_counter, lane, nullptr); break;
                                case 2: acc += lane_tasks<2>(M2, pitch2, (unsigned)c << sl, need, sl, &s_counter, lane, nullptr); break;
                                case 3: acc += lane_tasks<3>(M2, pitch2, (unsigned)c << sl, need, sl, &s_counter, lane, nullptr); break;
                                case 4: acc += lane_tasks<4>(M2, pitch2, (unsigned)c << sl, need, sl, &s_counter, lane, nullptr); break;
                                case 5: acc += lane_tasks<5>(M2, pitch2, (unsigned)c << sl, need, sl, &s_counter, lane, nullptr); break;
                                case 6: acc += lane_tasks<6>(M2, pitch2, (unsigned)c << sl, need, sl, &s_counter, lane, nullptr); break;
                                case 7: acc += lane_tasks<7>(M2, pitch2, (unsigned)c << sl, need, sl, &s_counter, lane, nullptr); break;
                                default: acc += lane_tasks<8>(M2, pitch2, (unsigned)c << sl, need, sl, &s_counter, lane, nullptr); break;
                            }
                        } else {
                            expand = true;
                        }
                    }
                    if (!expand) {
                        if (level == 0) break;
                        --level;
                        continue;
                    }
                    cursor[level] = -1;
                }
                // next member of this level's set after the cursor (every thread computes the same answer)
                int v = -1;
                {
                    const int from = cursor[level] + 1;
                    int w = from >> 6;
                    if (w < W1) {
                        u64 m = set[w] & (~0ull << (from & 63));
                        while (!m && ++w < W1) m = set[w];
                        if (m) v = (w << 6) + __ffsll((long long)m) - 1;
                    }
                }
                if (v < 0) {
                    if (level == 0) break;
                    --level;
                    continue;
                }
                cursor[level] = v;
                __syncthreads();          // everyone is done with the deeper levels before they are overwritten
                u64 *child = stack + (size_t)(level + 1) * P1;
                for (int w = tid; w < P1; w += BLOCK) child[w] = set[w] & M1[(size_t)v * P1 + w];
                __syncthreads();
                ++level;
                fresh = true;
            }
        }
    }
    const unsigned long long s = block_sum(acc, red);
    if (tid == 0 && s) atomicAdd(total, s);
}

}  // namespace lane
}  // namespace gmsb
