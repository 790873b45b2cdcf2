// CPU check of gms_b200/csrc/kclique_lane_core.cuh: the per-lane clique search, the residue-class task split and the
// set compaction are __host__ __device__, so exactly the code the GPU lanes run is driven here, task by task, against
// a plain recursive count on random DAG bit matrices.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../gms_b200/csrc/kclique_lane_core.cuh"

using namespace gmsb::lane;

static std::vector<std::vector<char>> random_dag(int c, double p, std::mt19937 &rng, bool ramp) {
    std::vector<std::vector<char>> adj(c, std::vector<char>(c, 0));
    std::uniform_real_distribution<double> u(0, 1);
    for (int a = 0; a < c; ++a)
        for (int b = a + 1; b < c; ++b) {
            const double q = ramp ? p * (0.3 + 0.7 * (double)(a + b) / (2.0 * c)) : p;     // denser towards the end
            adj[a][b] = u(rng) < q;
        }
    return adj;
}

// number of `need`-cliques whose members all lie in cand (ascending picks)
static u64 brute(const std::vector<std::vector<char>> &adj, const std::vector<int> &cand, int need) {
    if (need == 1) return cand.size();
    u64 t = 0;
    for (size_t i = 0; i < cand.size(); ++i) {
        std::vector<int> nxt;
        for (size_t j = i + 1; j < cand.size(); ++j)
            if (adj[cand[i]][cand[j]]) nxt.push_back(cand[j]);
        if ((int)nxt.size() >= need - 1) t += brute(adj, nxt, need - 1);
    }
    return t;
}

template <int NW>
static u64 run_lanes(const std::vector<u64> &cm, int pitch, int c, int need, int split_log2) {
    u64 total = 0;
    const unsigned ntasks = (unsigned)c << split_log2;
    for (unsigned t = 0; t < ntasks; ++t) total += lane_run_task<NW>(cm.data(), pitch, need, t, split_log2);
    return total;
}

// The flat form (FlatState / flat_refill / flat_step) driven the way kclique_lane.cuh: lane_tasks_flat drives it: 32 lane
// states, refills batched by flat_should_refill, one leaf operation per lane and iteration.
template <int NW>
static u64 run_flat(const std::vector<u64> &cm, int pitch, int c, int need, int split_log2) {
    FlatState<NW> s[32];
    for (auto &x : s) flat_init<NW>(x);
    unsigned counter = 0;
    const unsigned t_end = (unsigned)c << split_log2;
    u64 total = 0;
    for (;;) {
        int want = 0, live = 0;
        bool dry[32];
        for (int l = 0; l < 32; ++l) {
            dry[l] = s[l].w1 >= NW;
            want += dry[l] && !s[l].exhausted;
            live += !dry[l];
        }
        if (flat_should_refill(want, live)) {
            for (int l = 0; l < 32; ++l)
                if (dry[l] && !s[l].exhausted) {
                    flat_refill<NW>(s[l], cm.data(), pitch, need, split_log2, t_end, &counter, nullptr);
                    flat_skip<NW>(s[l]);
                }
            continue;
        }
        if (live == 0) break;
        for (int l = 0; l < 32; ++l)
            for (int r = 0; r < 4; ++r)                       // kFlatBurst
                if (s[l].w1 < NW) {
                    total += flat_step<NW>(s[l], cm.data(), pitch);
                    if (s[l].bits == 0) flat_skip<NW>(s[l]);
                }
    }
    return total;
}
static u64 run_flat_any(int nwb, const std::vector<u64> &cm, int pitch, int c, int need, int sl) {
    switch (nwb) {
        case 1: return run_flat<1>(cm, pitch, c, need, sl);
        case 2: return run_flat<2>(cm, pitch, c, need, sl);
        case 3: return run_flat<3>(cm, pitch, c, need, sl);
        default: return run_flat<4>(cm, pitch, c, need, sl);
    }
}

static u64 run_any(int nwb, const std::vector<u64> &cm, int pitch, int c, int need, int sl) {
    switch (nwb) {
        case 1: return run_lanes<1>(cm, pitch, c, need, sl);
        case 2: return run_lanes<2>(cm, pitch, c, need, sl);
        case 3: return run_lanes<3>(cm, pitch, c, need, sl);
        case 4: return run_lanes<4>(cm, pitch, c, need, sl);
        case 5: return run_lanes<5>(cm, pitch, c, need, sl);
        case 6: return run_lanes<6>(cm, pitch, c, need, sl);
        case 7: return run_lanes<7>(cm, pitch, c, need, sl);
        default: return run_lanes<8>(cm, pitch, c, need, sl);
    }
}

static int bucket(int c) { return (c + 63) >> 6; }

static u64 run_task_any(int nw, const std::vector<u64> &cm, int pitch, int need, unsigned task, int sl) {
    switch (nw) {
        case 1: return lane_run_task<1>(cm.data(), pitch, need, task, sl);
        case 2: return lane_run_task<2>(cm.data(), pitch, need, task, sl);
        case 3: return lane_run_task<3>(cm.data(), pitch, need, task, sl);
        case 4: return lane_run_task<4>(cm.data(), pitch, need, task, sl);
        case 5: return lane_run_task<5>(cm.data(), pitch, need, task, sl);
        case 6: return lane_run_task<6>(cm.data(), pitch, need, task, sl);
        case 7: return lane_run_task<7>(cm.data(), pitch, need, task, sl);
        default: return lane_run_task<8>(cm.data(), pitch, need, task, sl);
    }
}

int main() {
    std::mt19937 rng(12345);
    int fails = 0, checks = 0;
    // 1. lane search on compact matrices
    const int sizes[] = {1, 2, 3, 5, 31, 64, 65, 100, 128, 129, 200, 256, 300, 512};
    for (int c : sizes) {
        for (int variant = 0; variant < 3; ++variant) {
            const double p = variant == 0 ? 0.5 : variant == 1 ? 0.15 : 1.0;
            if (p == 1.0 && c > 40) continue;            // complete graphs: brute force explodes
            auto adj = random_dag(c, p, rng, variant == 0);
            const int nwb = bucket(c), pitch = pitch_for(nwb);
            std::vector<u64> cm((size_t)c * pitch, 0);
            for (int a = 0; a < c; ++a)
                for (int b = a + 1; b < c; ++b)
                    if (adj[a][b]) cm[(size_t)a * pitch + (b >> 6)] |= 1ull << (b & 63);
            std::vector<int> all(c);
            for (int i = 0; i < c; ++i) all[i] = i;
            for (int need = 3; need <= kMaxNeed; ++need) {
                if (c > 128 && p > 0.3 && need > 5) continue;        // keep the brute force fast
                if (c > 256 && need > 4 && p > 0.3) continue;
                const u64 want = brute(adj, all, need);
                for (int sl : {0, 3, 6}) {
                    const u64 got = run_any(nwb, cm, pitch, c, need, sl);
                    ++checks;
                    if (got != want) {
                        ++fails;
                        std::printf("FAIL lane c=%d p=%.2f need=%d split=%d got=%llu want=%llu\n", c, p, need, sl, got, want);
                    }
                    if (nwb <= 4) {
                        const u64 gf = run_flat_any(nwb, cm, pitch, c, need, sl);
                        ++checks;
                        if (gf != want) {
                            ++fails;
                            std::printf("FAIL flat c=%d p=%.2f need=%d split=%d got=%llu want=%llu\n", c, p, need, sl, gf, want);
                        }
                    }
                }
            }
        }
    }
    // 2. compaction: candidate set of a big matrix -> compact matrix, then the same count
    for (int D : {70, 300, 700, 1049}) {
        auto adj = random_dag(D, 0.5, rng, true);
        const int W1 = (D + 63) >> 6, P1 = W1 | 1;
        std::vector<u64> M1((size_t)D * P1, 0);
        for (int a = 0; a < D; ++a)
            for (int b = a + 1; b < D; ++b)
                if (adj[a][b]) M1[(size_t)a * P1 + (b >> 6)] |= 1ull << (b & 63);
        for (int trial = 0; trial < 6; ++trial) {
            // set = row i, optionally ANDed with another row, capped at kCMax members
            std::vector<u64> set(P1, 0);
            const int i = (int)(rng() % (unsigned)std::max(1, D / 3));
            for (int w = 0; w < P1; ++w) set[w] = M1[(size_t)i * P1 + w];
            if (trial & 1) {
                const int j = i + 1 + (int)(rng() % 5u);
                if (j < D) for (int w = 0; w < P1; ++w) set[w] &= M1[(size_t)j * P1 + w];
            }
            std::vector<int> members;
            for (int p = 0; p < D; ++p)
                if ((set[p >> 6] >> (p & 63)) & 1) {
                    if ((int)members.size() == kCMax) set[p >> 6] &= ~(1ull << (p & 63));
                    else members.push_back(p);
                }
            const int c = (int)members.size();
            if (c < 3) continue;
            std::vector<int> prefix(P1 + 1, 0);
            for (int w = 0; w < P1; ++w) prefix[w + 1] = prefix[w] + popc64(set[w]);
            std::vector<int> list(c);
            for (int p = 0; p < D; ++p)
                if ((set[p >> 6] >> (p & 63)) & 1) list[compact_index(set.data(), prefix.data(), p)] = p;
            for (int a = 0; a < c; ++a)
                if (list[a] != members[a]) { ++fails; std::printf("FAIL list D=%d a=%d\n", D, a); break; }
            const int nwb = bucket(c), pitch = pitch_for(nwb);
            std::vector<u64> M2((size_t)c * pitch, ~0ull);
            for (int a = 0; a < c; ++a)
                compact_row(set.data(), prefix.data(), W1, M1.data() + (size_t)list[a] * P1, list[a],
                            M2.data() + (size_t)a * pitch, nwb);
            bool ok = true;
            for (int a = 0; a < c && ok; ++a)
                for (int b = 0; b < 64 * nwb && ok; ++b) {
                    const bool got = (M2[(size_t)a * pitch + (b >> 6)] >> (b & 63)) & 1;
                    const bool want = b < c && b > a && adj[list[a]][list[b]];
                    if (got != want) { ok = false; ++fails; std::printf("FAIL compact D=%d a=%d b=%d\n", D, a, b); }
                }
            ++checks;
            for (int need : {3, 4}) {
                const u64 want = brute(adj, members, need);
                const u64 got = run_any(nwb, M2, pitch, c, need, 4);
                ++checks;
                if (got != want) { ++fails; std::printf("FAIL compact-count D=%d c=%d need=%d got=%llu want=%llu\n", D, c, need, got, want); }
            }
        }
    }
    // 3. third level (kclique_lane.cuh: warp_tasks): below a member a whose row has <= 128 members the search runs in
    //    a re-indexed matrix M3 (pitch 3) with need-1; larger rows are searched in cm as 64 residue-class tasks
    for (int c : {90, 200, 220, 400, 512}) {
        auto adj = c == 220 ? random_dag(c, 0.95, rng, false)         // rows of up to ~210 members: too large for M3
                            : random_dag(c, c <= 200 ? 0.6 : 0.3, rng, true);
        const int nw = bucket(c), pitch = pitch_for(nw);
        std::vector<u64> cm((size_t)c * pitch, 0);
        for (int a = 0; a < c; ++a)
            for (int b = a + 1; b < c; ++b)
                if (adj[a][b]) cm[(size_t)a * pitch + (b >> 6)] |= 1ull << (b & 63);
        std::vector<int> all(c);
        for (int i = 0; i < c; ++i) all[i] = i;
        for (int need : {5, 6}) {
            if (c > 200 && need > 5) continue;
            const u64 want = brute(adj, all, need);
            u64 got = 0;
            int compacted = 0, big = 0;
            for (int a = 0; a < c; ++a) {
                const u64 *row = cm.data() + (size_t)a * pitch;
                int prefix[8] = {0}, c3 = 0;
                for (int w = 0; w < nw; ++w) { prefix[w] = c3; c3 += popc64(row[w]); }
                if (c3 < need - 1) continue;
                if (c3 > kC3Max) {
                    ++big;
                    for (unsigned st = 0; st < 64; ++st) got += run_task_any(nw, cm, pitch, need, ((unsigned)a << 6) | st, 6);
                    continue;
                }
                ++compacted;
                std::vector<int> list(c3);
                for (int p = 0; p < nw * 64; ++p)
                    if ((row[p >> 6] >> (p & 63)) & 1) list[compact_index(row, prefix, p)] = p;
                const int nw3 = (c3 + 63) >> 6;
                std::vector<u64> m3((size_t)c3 * kP3, ~0ull);
                for (int m = 0; m < c3; ++m)
                    compact_row(row, prefix, nw, cm.data() + (size_t)list[m] * pitch, list[m], m3.data() + (size_t)m * kP3, nw3);
                got += run_any(nw3, m3, kP3, c3, need - 1, 3);
            }
            ++checks;
            if (got != want) { ++fails; std::printf("FAIL third-level c=%d need=%d got=%llu want=%llu\n", c, need, got, want); }
            std::printf("third level c=%d need=%d: %d rows re-indexed, %d searched in place\n", c, need, compacted, big);
        }
    }
    std::printf("%d checks, %d failures\n", checks, fails);
    if (!fails) std::printf("all checks passed\n");
    return fails ? 1 : 0;
}
