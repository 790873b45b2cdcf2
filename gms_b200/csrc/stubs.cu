// stubs.cu — operators not implemented yet fail loudly (never a CPU fallback).
#include "common.cuh"
#include "ops.cuh"
namespace gmsb {
void tc_vertex2(Graph &, int64_t *) { throw Error(GMSB_ERR_UNSUPPORTED, "tc_vertex2: not implemented yet"); }
void degeneracy_rank(Graph &, vid_t *) { throw Error(GMSB_ERR_UNSUPPORTED, "order_degeneracy: not implemented yet"); }
void kclique_count(Graph &, int, uint64_t *) { throw Error(GMSB_ERR_UNSUPPORTED, "kclique_count: not implemented yet"); }
void kclique_count_ordered(Graph &, int, uint64_t *) { throw Error(GMSB_ERR_UNSUPPORTED, "kclique_count_ordered: not implemented yet"); }
}  // namespace gmsb
