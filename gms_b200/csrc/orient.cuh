// orient.cuh — the rank-space oriented DAG and the entry points that build it.
#pragma once
#include "common.cuh"

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
    TcPlan *plan = nullptr;      // cached schedule for the triangle kernels
    ~Dag();
};

void degree_order(const Graph &g, DevBuf<vid_t> &order, DevBuf<vid_t> &rank, int64_t *max_deg = nullptr);
void orient_by_rank(const Graph &g, const vid_t *rank_dev, DevBuf<eid_t> &doff, DevBuf<vid_t> &dnbr, int64_t *m_out,
                    int *max_dplus, DevBuf<int32_t> *dplus = nullptr);
Dag *build_degree_dag(const Graph &g);
// Pipelined form used while a host CSR is still being uploaded (graph_build.cu): the ranking needs the offsets only,
// and the relabel + count pass runs per vertex range as soon as that range's neighbour slots have arrived.
struct OrientPipeline {
    Dag *d = nullptr;
    DevBuf<vid_t> rnbr, bigq;
    DevBuf<int> nbigq;
};
void orient_pipeline_begin(const Graph &g, OrientPipeline &p);                       // offsets on the device
void orient_pipeline_range(const Graph &g, OrientPipeline &p, int64_t u_begin, int64_t u_end);   // slots of the range too
void orient_pipeline_finish(const Graph &g, OrientPipeline &p);                      // everything enqueued: emit + sorts
Graph *induce_directed(const Graph &g, const vid_t *ranking_host);

// graph_build.cu
Graph *graph_from_csr_device(int64_t n, const eid_t *off, const vid_t *nbr, bool directed, bool host_src);
Graph *graph_from_csr_host_pipelined(int64_t n, const eid_t *off, const vid_t *nbr);
Graph *graph_from_edgelist_device(int64_t m, const vid_t *src, const vid_t *dst, bool symmetrize);
Graph *graph_relabel_by_degree(const Graph &in);

// tc.cu
void tc_total(Graph &g, const gmsb_tc_options &opt, uint64_t *out, gmsb_tc_stats *stats);

}  // namespace gmsb
