// graph_io_test <undirected.sg> <directed.sg> <outdir>: load with the C++ reader, count triangles on the device,
// write both back (the Python test compares the bytes with the reference-written originals).
#include <cstdio>
#include <gms_b200/gms_api.hpp>
#include <gms_b200/graph_io.hpp>

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    try {
        CudaSetGraph g = gms_b200::LoadGraph(argv[1]);
        std::printf("triangles %zu\n", GMS::TriangleCount::Par::count_total<CudaSetGraph>(g));
        gms_b200::WriteSerializedGraph(g, std::string(argv[3]) + "/out.sg");
        CudaSetGraph d = gms_b200::LoadGraph(argv[2]);
        if (!d.directed()) return 1;
        gms_b200::WriteSerializedGraph(d, std::string(argv[3]) + "/dir.sg");
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 3;
    }
    return 0;
}
