// graph_io.hpp — graph files in the reference's formats, for CudaSetGraph (host side).
//
//   .el  text edge list "u v" per line                  gms/third_party/gapbs/reader.h:58-72, writer.h:32-37
//   .sg  serialized CSR: bool directed, int64 CSR slots, int64 num_nodes, int64 offsets[n+1], int32 neighbours[];
//        a directed graph appends the inverse offsets + neighbours
//                                                        gms/third_party/gapbs/writer.h:39-70, reader.h:252-305
//
// LoadGraph mirrors Builder::MakeGraph for files (gapbs/builder.h:1642-1660): a .sg file is used as is, a .el file
// goes through MakeGraphFromEL + SquishGraph — here on the GPU (CudaSetGraph::FromEdgeList).
#pragma once
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

#include "cuda_set_graph.hpp"

namespace gms_b200 {

namespace detail {
inline bool ends_with(const std::string &s, const std::string &suffix) {
    return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}
}  // namespace detail

inline void ReadEdgeList(const std::string &path, std::vector<NodeId> &src, std::vector<NodeId> &dst) {
    std::ifstream in(path);
    if (!in.is_open()) throw std::runtime_error("Couldn't open file " + path);      // reader.h:226-229
    long long u, v;
    while (in >> u >> v) { src.push_back(static_cast<NodeId>(u)); dst.push_back(static_cast<NodeId>(v)); }
}

inline CudaSetGraph ReadSerializedGraph(const std::string &path) {
    std::ifstream in(path, std::ios::binary);
    if (!in.is_open()) throw std::runtime_error("Couldn't open file " + path);      // reader.h:271-274
    bool directed = false;
    int64_t slots = 0, n = 0;
    in.read(reinterpret_cast<char *>(&directed), sizeof(bool));
    in.read(reinterpret_cast<char *>(&slots), sizeof(int64_t));
    in.read(reinterpret_cast<char *>(&n), sizeof(int64_t));
    if (!in || n < 0 || slots < 0) throw std::runtime_error(path + ": not a .sg file");
    std::vector<int64_t> off(static_cast<size_t>(n) + 1);
    std::vector<NodeId> nbr(static_cast<size_t>(slots));
    in.read(reinterpret_cast<char *>(off.data()), static_cast<std::streamsize>(off.size() * sizeof(int64_t)));
    in.read(reinterpret_cast<char *>(nbr.data()), static_cast<std::streamsize>(nbr.size() * sizeof(NodeId)));
    if (!in || off[static_cast<size_t>(n)] != slots) throw std::runtime_error(path + ": truncated .sg file");
    return CudaSetGraph::FromCSR(n, off.data(), nbr.data(), directed);
}

inline void WriteSerializedGraph(const CudaSetGraph &g, const std::string &path) {
    const int64_t n = g.num_nodes(), slots = g.num_edges_directed();
    std::vector<int64_t> off(static_cast<size_t>(n) + 1, 0);
    std::vector<NodeId> nbr(static_cast<size_t>(slots));
    for (int64_t u = 0; u < n; ++u) {
        int64_t p = off[u];
        for (NodeId v : g.out_neigh(static_cast<NodeId>(u))) nbr[p++] = v;
        off[u + 1] = p;
    }
    std::ofstream out(path, std::ios::binary);
    if (!out.is_open()) throw std::runtime_error("Couldn't write to file " + path);
    const bool directed = g.directed();
    out.write(reinterpret_cast<const char *>(&directed), sizeof(bool));
    out.write(reinterpret_cast<const char *>(&slots), sizeof(int64_t));
    out.write(reinterpret_cast<const char *>(&n), sizeof(int64_t));
    out.write(reinterpret_cast<const char *>(off.data()), static_cast<std::streamsize>(off.size() * sizeof(int64_t)));
    out.write(reinterpret_cast<const char *>(nbr.data()), static_cast<std::streamsize>(nbr.size() * sizeof(NodeId)));
    if (directed) {      // the inverse, ascending lists (writer.h:64-68)
        std::vector<int64_t> ioff(static_cast<size_t>(n) + 1, 0);
        for (NodeId v : nbr) ioff[static_cast<size_t>(v) + 1]++;
        for (int64_t i = 0; i < n; ++i) ioff[i + 1] += ioff[i];
        std::vector<NodeId> inbr(static_cast<size_t>(slots));
        std::vector<int64_t> cur(ioff.begin(), ioff.end() - 1);
        for (int64_t u = 0; u < n; ++u)
            for (int64_t e = off[u]; e < off[u + 1]; ++e) inbr[cur[nbr[e]]++] = static_cast<NodeId>(u);
        out.write(reinterpret_cast<const char *>(ioff.data()), static_cast<std::streamsize>(ioff.size() * sizeof(int64_t)));
        out.write(reinterpret_cast<const char *>(inbr.data()), static_cast<std::streamsize>(inbr.size() * sizeof(NodeId)));
    }
}

inline void WriteEdgeList(const CudaSetGraph &g, const std::string &path) {
    std::ofstream out(path);
    if (!out.is_open()) throw std::runtime_error("Couldn't write to file " + path);
    for (int64_t u = 0; u < g.num_nodes(); ++u)
        for (NodeId v : g.out_neigh(static_cast<NodeId>(u))) out << u << " " << v << "\n";
}

inline CudaSetGraph LoadGraph(const std::string &path, bool symmetrize = true) {
    if (detail::ends_with(path, ".sg")) return ReadSerializedGraph(path);
    if (detail::ends_with(path, ".el")) {
        std::vector<NodeId> s, d;
        ReadEdgeList(path, s, d);
        return CudaSetGraph::FromEdgeList(s.data(), d.data(), static_cast<int64_t>(s.size()), symmetrize);
    }
    throw std::runtime_error("Unrecognized suffix: " + path);                         // reader.h:243-245
}

}  // namespace gms_b200
