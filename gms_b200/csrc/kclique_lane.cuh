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
//     deals its tasks to the lanes; with two vertices left the pairs are counted straight off the big matrix.
// A lane keeps its candidate sets in registers (kclique_lane_core.cuh); only the path is remembered per level, the
// parent's set is recomputed on the way back, so there is no per-lane stack memory at all.
#pragma once
#include "common.cuh"
#include "isect.cuh"
#include "kclique_lane_core.cuh"

namespace gmsb {
namespace lane {

// count of `need`-cliques (need >= 3) in the compact graph cm[c][pitch]; tasks come from *counter (zeroed by the
// caller, one per call); all 32 lanes of every calling warp must be here.  Returns this lane's partial count.
template <int NW>
__device__ u64 lane_tasks(const u64 *cm, int pitch, int c, int need, int split_log2, unsigned *counter, int lane) {
    const unsigned ntasks = (unsigned)c << split_log2;
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
                if (t >= ntasks) exhausted = true;
                else lane_begin<NW>(s, cm, pitch, t, split_log2);
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

__device__ __forceinline__ int split_for(int c, int block) {
    int sl = 0;
    while (sl < 6 && (c << sl) < 16 * block) ++sl;       // >= 16 tasks per lane keeps the tail short
    return sl;
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
                   unsigned int *__restrict__ ticket, int pi, int P) {
    constexpr int DMAX = 64 * NWB;
    constexpr int PITCH = pitch_for(NWB);
    __shared__ vid_t S[DMAX];
    __shared__ u64 M[DMAX * PITCH];
    __shared__ unsigned long long red[BLOCK / 32];
    __shared__ unsigned int s_item, s_counter;
    const int tid = threadIdx.x, lane = tid & 31;
    u64 acc = 0;
    for (;;) {
        __syncthreads();
        if (tid == 0) { s_item = atomicAdd(ticket, 1u); s_counter = 0; }
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
        acc += lane_tasks<NWB>(M, PITCH, D, k - 1, split_for(D, BLOCK), &s_counter, lane);
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
                                case 1: acc += lane_tasks<1>(M2, pitch2, c, need, sl, &s_counter, lane); break;
                                case 2: acc += lane_tasks<2>(M2, pitch2, c, need, sl, &s_counter, lane); break;
                                case 3: acc += lane_tasks<3>(M2, pitch2, c, need, sl, &s_counter, lane); break;
                                case 4: acc += lane_tasks<4>(M2, pitch2, c, need, sl, &s_counter, lane); break;
                                case 5: acc += lane_tasks<5>(M2, pitch2, c, need, sl, &s_counter, lane); break;
                                case 6: acc += lane_tasks<6>(M2, pitch2, c, need, sl, &s_counter, lane); break;
                                case 7: acc += lane_tasks<7>(M2, pitch2, c, need, sl, &s_counter, lane); break;
                                default: acc += lane_tasks<8>(M2, pitch2, c, need, sl, &s_counter, lane); break;
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
