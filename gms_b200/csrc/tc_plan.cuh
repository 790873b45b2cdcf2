// tc_plan.cuh — the triangle schedule ("plan") shared by the counting kernels (tc.cu) and the per-edge support
// kernels (tc_support.cu): suffix descriptors grouped by their closing vertex v, hub items, light-edge lists.
// The grouping is a counting sort keyed on (v, length class): one pass counts, a scan lays the segments out, a second
// pass scatters every descriptor to its segment through an atomic cursor (tc.cu: build_plan).
#pragma once
#include "common.cuh"
#include "orient.cuh"
#include "owner.cuh"

namespace gmsb {

constexpr int kLenBits = 24;                        // descriptor = (start << 24) | len
constexpr uint64_t kLenMask = (1ull << kLenBits) - 1;
// Descriptors of one v are ordered by length class so that short suffixes can share a warp:
constexpr int kClassBits = 2;                       // sort key = (v << 2) | class
constexpr int kShortLen = 8;                        // class 0: len <= 8   -> one lane per descriptor
constexpr int kMidLen = 96;                         // class 1: len <= 96  -> eight lanes per descriptor
                                                    // class 2: longer     -> one warp per descriptor

constexpr int kSmallWindowBytes = 55000;            // 4 x (55 KB bitmap + static) <= 227 KB per SM

struct Item {            // one CTA work unit: a slice of hub v's incoming descriptors
    int32_t v;
    int32_t count;       // descriptors in the slice
    int32_t n0, n1;      // [0,n0) class 0, [n0,n1) class 1, [n1,count) class 2
    int64_t begin;
    int32_t pad[2];
};

struct TcPlan {
    gmsb_tc_options opt{};
    int64_t n_desc = 0;                 // descriptors that can close a triangle
    DevBuf<Item> items;                 // bitmap work items
    int64_t n_items = 0;
    int max_span_words = 0;
    // items are ordered by window class: NEAR (the whole range after v up to n-1 fits the small window: no clamp in
    // the probe), small window (kSmallWindowBytes of bitmap: four CTAs per SM, 64 resident warps), wide (three per SM)
    int64_t cls_items[3] = {0, 0, 0};
    int cls_words[3] = {0, 0, 0};
    DevBuf<uint64_t> m_desc, g_desc;    // light edges for merge / gallop
    DevBuf<vid_t> m_v, g_v;
    int64_t n_merge = 0, n_gallop = 0, n_bitmap_edges = 0;
    // (with part_count > 1 every figure below covers this device's share of the edges: those into the vertices it owns)
    uint64_t algorithmic_bytes = 0;     // B_TC over the oriented edges
    uint64_t wedges = 0;
    uint64_t bytes_bitmap = 0, bytes_kept = 0, wedges_bitmap = 0;
    int max_hub_dplus = 0;              // largest d+ among hub vertices (tc_support sizes its shared lists with it)
    // hub descriptors grouped by v (ascending), then length class; the order inside a class is the order in which the
    // scatter pass reached the edges (roughly ascending u)
    DevBuf<uint64_t> desc;
};

// Builds (or reuses) the degree-oriented DAG and the schedule cached on the graph handle.
TcPlan &ensure_plan(Graph &g, const gmsb_tc_options &opt_normalised);
gmsb_tc_options normalise_tc_options(const gmsb_tc_options *in);

__device__ __forceinline__ uint32_t probe(const uint32_t *bm, uint32_t x, uint32_t cap_words) {
    return (bm[min(x >> 5, cap_words)] >> (x & 31)) & 1u;
}

}  // namespace gmsb
