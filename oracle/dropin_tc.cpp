// TEST INFRASTRUCTURE ONLY — the drop-in demonstration.
//
// A benchmark main in the shape of the reference's own triangle-count mains, but with gms_b200::CudaSetGraph as the
// graph type.  Everything it calls besides CudaSetGraph is the UNMODIFIED reference, compiled from where it lies by
// oracle/Makefile into oracle/_ref/dropin_tc: the CLI (parse_and_load: generator, builder, auto-relabel), the
// harness (BenchmarkKernelBk: FromCGraph once, kernel per trial, "@@@" result lines) and the -v verifiers
// (Verify::total_count, Verify::vertex_count<2>), which therefore check the GPU results at run time.
//
//   oracle/_ref/dropin_tc -g kronecker 16 --deg 16 -n 3 -v
//   GMSB_DEVICES=0,1,2,3,4,5,6,7 oracle/_ref/dropin_tc -g kronecker 20 --deg 16 -n 3 -v      (one process, eight GPUs)
#include <gms/third_party/gapbs/benchmark.h>
#include <gms/common/cli/cli.h>
#include <gms/representations/graphs/set_graph.h>
#include <gms/common/benchmark.h>
#include <gms/algorithms/set_based/triangle_count/triangle_count.h>
#include <gms/algorithms/set_based/triangle_count/verifier.h>
#include <gms/algorithms/preprocessing/preprocessing.h>

#define GMSB_WITH_GMS_HEADERS
#include <gms_b200/gms_api.hpp>

namespace tc = GMS::TriangleCount;
using GMS::BenchmarkKernelBk;
using gms_b200::CudaSetGraph;

int main(int argc, char *argv[]) {
    if (gms_b200::UseDevicesFromEnv()) std::cout << "gms-b200: " << gms_b200::multi_device_count() << " devices" << std::endl;
    auto loaded = GMS::CLI::Parser().parse_and_load(argc, argv);
    GMS::CLI::Args &args = std::get<0>(loaded);
    CSRGraph &host_graph = std::get<1>(loaded);

    // total count: the B200 path behind the reference's harness and verifier
    BenchmarkKernelBk<CudaSetGraph>(args, host_graph, tc::Par::count_total<CudaSetGraph>, tc::Verify::total_count,
                                    "tc-total-par-CudaSetGraph");

    // per-vertex counts: the harness wants a kernel that RETURNS its result
    auto per_vertex = [](const CudaSetGraph &g) {
        std::vector<int64_t> counts;
        tc::Par::vertex_count2<CudaSetGraph, std::vector<int64_t>>(g, counts);
        return counts;
    };
    BenchmarkKernelBk<CudaSetGraph>(args, host_graph, per_vertex, tc::Verify::vertex_count<2>,
                                    "tc-vertex-count2-par-CudaSetGraph");

    // the reference's own SortedSet path on the same graph, for the side-by-side "@@@" line (small graphs only)
    if (host_graph.num_nodes() <= (1 << 18))
        BenchmarkKernelBk<SortedSetGraph>(args, host_graph, tc::Par::count_total<SortedSetGraph>,
                                          tc::Verify::total_count, "tc-total-par-SortedSetGraph");
    return 0;
}
