// examples/triangle_counting.cpp — the stand-alone counterpart of the reference's examples/triangle_counting.cpp:
// same shape (load or generate a graph, build the set graph once, time `kernel(graph)` for a few trials, verify,
// print an "@@@" line per trial as gms/common/benchmark.h:129-133 does), with CudaSetGraph as the graph type.
//
//   g++ -std=c++17 -O2 -Iinclude examples/triangle_counting.cpp -Lgms_b200/lib -lgmsb -Wl,-rpath,$PWD/gms_b200/lib -o tc
//   ./tc -g 20 [--deg 16] [-n 3] [-v] [-k 4]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include <gms_b200/gms_api.hpp>

using namespace GMS;

static double seconds_since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int main(int argc, char *argv[]) {
    int scale = 16, degree = 16, trials = 3, clique = 0;
    bool verify = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "-g" && i + 1 < argc) scale = std::atoi(argv[++i]);
        else if (a == "--deg" && i + 1 < argc) degree = std::atoi(argv[++i]);
        else if (a == "-n" && i + 1 < argc) trials = std::atoi(argv[++i]);
        else if (a == "-k" && i + 1 < argc) clique = std::atoi(argv[++i]);
        else if (a == "-v") verify = true;
        else { std::fprintf(stderr, "usage: %s -g <scale> [--deg d] [-n trials] [-v] [-k clique-size]\n", argv[0]); return 2; }
    }
    try {
        auto t0 = std::chrono::steady_clock::now();
        CudaSetGraph graph = CudaSetGraph::Kronecker(scale, degree);
        std::printf("%-21s%3.5lf\n", "GraphExec buildTime:", seconds_since(t0));
        std::printf("Graph has %lld nodes and %lld undirected edges\n", (long long)graph.num_nodes(),
                    (long long)graph.num_edges());
        for (int t = 0; t < trials; ++t) {
            t0 = std::chrono::steady_clock::now();
            size_t result = TriangleCount::Par::count_total<CudaSetGraph>(graph);
            double trial = seconds_since(t0);
            std::printf("%-21s%3.5lf\n", "Trial Time:", trial);
            if (verify) {
                // independent check: per-vertex counts come from a different kernel family (support kernels)
                t0 = std::chrono::steady_clock::now();
                std::vector<int64_t> counts;
                TriangleCount::Par::vertex_count2<CudaSetGraph>(graph, counts);
                long long sum = 0;
                for (int64_t c : counts) sum += c;
                bool pass = (size_t)(sum / 6) == result && sum % 6 == 0;
                std::printf("@@@ %g %s %g tc-total-par-CudaSetGraph triangles=%zu\n", trial, pass ? "PASS" : "FAIL",
                            seconds_since(t0), result);
                if (!pass) return 1;
            } else {
                std::printf("@@@ %g tc-total-par-CudaSetGraph triangles=%zu\n", trial, result);
            }
        }
        if (clique > 0) {
            std::vector<NodeId> ranking;
            PpParallel::getDegreeOrdering<CudaSetGraph, true>(graph, ranking);
            CudaSetGraph dag = PpSequential::InduceDirectedGraph(graph, ranking);
            t0 = std::chrono::steady_clock::now();
            unsigned long long c = KClique::Par::EP_kclisting(dag, KClique::CLCliqueApp(clique));
            std::printf("@@@ %g kclique-ep-CudaSetGraph k=%d count=%llu\n", seconds_since(t0), clique, c);
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 3;
    }
    return 0;
}
