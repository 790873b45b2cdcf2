// orient.cuh — the rank-space oriented DAG and the entry points that build it.
#pragma once
#include "common.cuh"
#include <vector>

namespace gmsb {

struct TcPlan;
void delete_plan(TcPlan *p);     // tc.cu

// Degree-oriented DAG of an undirected graph, in rank space (vertex id == position in (degree asc, id asc)).
struct Dag {
    int64_t n = 0;
    int64_t m = 0;               // oriented edges
    int max_dplus = 0;
    int64_t max_deg = 0;         // largest degree of the undirected graph
    DevBuf<vid_t> order;         // rank -> original id
    DevBuf<vid_t> rank;          // original id -> rank
    DevBuf<eid_t> off;           // n+1
    DevBuf<vid_t> nbr;           // m, ascending within each list, all entries > owner
    DevBuf<int32_t> dplus;       // n: d+(v) as a compact array (random reads of it stay inside the L2)
    DevBuf<uint32_t> spos;       // m, built on demand by the triangle schedule (tc.cu: ensure_slot_positions): per slot
                                 // (position in its list << 16) | elements after it
    TcPlan *plan = nullptr;      // cached schedule for the triangle kernels
    ~Dag();
};

void degree_order(const Graph &g, DevBuf<vid_t> &order, DevBuf<vid_t> &rank, int64_t *max_deg = nullptr);
void orient_by_rank(const Graph &g, const vid_t *rank_dev, DevBuf<eid_t> &doff, DevBuf<vid_t> &dnbr, int64_t *m_out,
                    int *max_dplus, DevBuf<int32_t> *dplus = nullptr);
Dag *build_degree_dag(const Graph &g);
// The orientation passes work on vertex ranges of the input and write each row's survivors to the front of the row's
// own slot range (orient.cu); this is their state between calls.
struct OrientRows {
    DevBuf<vid_t> rnbr, trow, bigq;     // relabelled slots; compacted + sorted rows in place; vertices with big lists
    DevBuf<int32_t> dold;               // d+ by original id
    DevBuf<uint64_t> queue;             // lists waiting for a sorter: (start << 24 | length)
    DevBuf<int> counters;
    int big_done = 0, sorted_mid = 0, sorted_long = 0;
};
void orient_rows_begin(const Graph &g, OrientRows &w);
void orient_rows_range(const Graph &g, const vid_t *rank_dev, OrientRows &w, int64_t u_begin, int64_t u_end);
void orient_rows_big(const Graph &g, const vid_t *rank_dev, OrientRows &w);
void orient_rows_sort(OrientRows &w);
// Pipelined form used while a host CSR is still being uploaded (graph_build.cu): the ranking needs the offsets only,
// and the relabel / emit / sort passes run per vertex range as soon as that range's neighbour slots have arrived.
struct OrientPipeline {
    Dag *d = nullptr;
    OrientRows w;
};
void orient_pipeline_begin(const Graph &g, OrientPipeline &p);                       // offsets on the device
void orient_pipeline_range(const Graph &g, OrientPipeline &p, int64_t u_begin, int64_t u_end);   // slots of the range too
void orient_pipeline_finish(const Graph &g, OrientPipeline &p);                      // big lists, offsets, rows into place
// Sharded form (one vertex range per device, graph_build.cu: shard_*)
int64_t orient_piece_layout(const Graph &g, const vid_t *rank_dev, OrientRows &w, int64_t u0, int64_t u1,
                            DevBuf<eid_t> &piece_off);
void orient_piece_export(const Graph &g, OrientRows &w, int64_t u0, int64_t u1, const DevBuf<eid_t> &piece_off,
                         vid_t *piece_dst, int32_t *dplus_all);
void dag_from_pieces(Dag &d, const int32_t *dplus_all, const vid_t *pieces, int64_t stride, const int64_t *cut, int parts);
Graph *induce_directed(const Graph &g, const vid_t *ranking_host);

// graph_build.cu
Graph *graph_from_csr_device(int64_t n, const eid_t *off, const vid_t *nbr, bool directed, bool host_src);
Graph *graph_from_csr_host_pipelined(int64_t n, const eid_t *off, const vid_t *nbr);
// sharded construction: see graph_build.cu
struct Shard {
    Graph *g = nullptr;
    Dag *d = nullptr;
    OrientRows w;
    int part = 0, parts = 1;
    std::vector<int64_t> cut;           // parts + 1 vertex cuts
    int64_t piece_len = 0;
    DevBuf<eid_t> piece_off;            // exclusive scan of d+ over the own range
    ~Shard();
};
Shard *shard_begin(int64_t n, const eid_t *off, const vid_t *nbr, const eid_t *off_dev, int part, int parts,
                   int64_t *piece_len);
void shard_export(Shard &s, vid_t *piece_dev, int32_t *dplus_all_dev);
Graph *shard_finish(Shard &s, const vid_t *pieces_dev, int64_t piece_stride, const int32_t *dplus_all_dev);
Graph *graph_from_edgelist_device(int64_t m, const vid_t *src, const vid_t *dst, bool symmetrize);
Graph *graph_relabel_by_degree(const Graph &in);

// tc.cu
void tc_total(Graph &g, const gmsb_tc_options &opt, uint64_t *out, gmsb_tc_stats *stats);

}  // namespace gmsb
