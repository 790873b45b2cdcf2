// TEST INFRASTRUCTURE ONLY — the drop-in demonstration: the reference's OWN benchmark main
// (examples/triangle_counting.cpp, gms/algorithms/set_based/triangle_count/triangle_count.cc) with the graph type
// swapped for gms_b200::CudaSetGraph.  Everything else — CLI parsing, generator, builder, auto-relabel,
// BenchmarkKernelBk, the -v verifier (Verify::total_count, Verify::vertex_count<2>) and the @@@ result lines —
// is the unmodified reference, compiled from where it lies by oracle/Makefile into oracle/_ref/dropin_tc.
//
//   oracle/_ref/dropin_tc -g kronecker 16 --deg 16 -n 3 -v
#include <gms/third_party/gapbs/benchmark.h>
#include <gms/common/cli/cli.h>
#include <gms/representations/graphs/set_graph.h>
#include <gms/common/benchmark.h>
#include <gms/algorithms/set_based/triangle_count/triangle_count.h>
#include <gms/algorithms/set_based/triangle_count/verifier.h>
#include <gms/algorithms/preprocessing/preprocessing.h>

#define GMSB_WITH_GMS_HEADERS
#include <gms_b200/gms_api.hpp>

using namespace GMS;
using namespace GMS::TriangleCount;
using gms_b200::CudaSetGraph;

template <class AnyGraph, class Fn>
constexpr auto output_wrap(Fn fn) {
    return [fn{std::move(fn)}](const AnyGraph &g) {
        std::vector<int64_t> output;
        fn(g, output);
        return output;
    };
}

int main(int argc, char *argv[]) {
    auto [args, g] = CLI::Parser().parse_and_load(argc, argv);

    // the B200 path behind the reference's harness
    BenchmarkKernelBk<CudaSetGraph>(args, g, Par::count_total<CudaSetGraph>, Verify::total_count,
                                    "tc-total-par-CudaSetGraph");
    BenchmarkKernelBk<CudaSetGraph>(args, g, output_wrap<CudaSetGraph>(Par::vertex_count2<CudaSetGraph, std::vector<int64_t>>),
                                    Verify::vertex_count<2>, "tc-vertex-count2-par-CudaSetGraph");
    // the reference's own SortedSet path, for the side-by-side @@@ line
    if (g.num_nodes() <= (1 << 18))
        BenchmarkKernelBk<SortedSetGraph>(args, g, Par::count_total<SortedSetGraph>, Verify::total_count,
                                          "tc-total-par-SortedSetGraph");
    return 0;
}
