// ops.cuh — host-callable operators implemented across the .cu files (all take HOST output buffers).
#pragma once
#include "common.cuh"

namespace gmsb {

// generator.cu (host side)
void generate_rmat(int scale, int64_t m, float a, float b, float c, bool permute, vid_t *src, vid_t *dst);
void generate_uniform(int scale, int64_t m, vid_t *src, vid_t *dst);

// setops.cu
void intersect_count_batch(Graph &g, int64_t np, const vid_t *a, const vid_t *b, uint64_t *out);
void intersect_batch(Graph &g, int64_t np, const vid_t *a, const vid_t *b, int64_t *out_offsets, vid_t *out_elems,
                     int64_t cap);
void difference_batch(Graph &g, int64_t np, const vid_t *a, const vid_t *b, int64_t *out_offsets, vid_t *out_elems,
                      int64_t cap);
void union_batch(Graph &g, int64_t np, const vid_t *a, const vid_t *b, int64_t *out_offsets, vid_t *out_elems,
                 int64_t cap);
void union_count_batch(Graph &g, int64_t np, const vid_t *a, const vid_t *b, uint64_t *out);
void pair_similarity(Graph &g, int metric, int64_t np, const vid_t *a, const vid_t *b, double *out);
void edge_similarity(Graph &g, int metric, double *out, int64_t *m_out);

// tc_support.cu
void tc_vertex2(Graph &g, int64_t *out_n);
void edge_scores_from_support(Graph &g, int metric, const int64_t *base_dev, double *out_dev);

// kcore.cu
void degeneracy_rank(Graph &g, vid_t *out_rank);
void degeneracy_order_approx(Graph &g, double epsilon, bool rank_format, vid_t *out_host);

// kclique.cu
void kclique_count(Graph &g, int k, uint64_t *out, int part_index = 0, int part_count = 1);
void kclique_count_ordered(Graph &g, int k, uint64_t *out);

}  // namespace gmsb
