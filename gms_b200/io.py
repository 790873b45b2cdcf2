"""Graph files in the reference's formats (host side).

  .el  text edge list, one "u v" pair per line                 gms/third_party/gapbs/reader.h:58-72, writer.h:32-37
  .sg  serialized CSR: bool directed, int64 edges_to_write (CSR slots), int64 num_nodes, int64 offsets[n+1],
       int32 neighbours[slots]; a directed graph appends the inverse offsets + neighbours
                                                                gms/third_party/gapbs/writer.h:39-70, reader.h:252-305

Reading produces host arrays; `load_graph` hands them to the device builders (the same route `Builder::MakeGraph`
takes: .sg is used as is, .el goes through MakeGraphFromEL + SquishGraph, here on the GPU).
"""
import numpy as np


def read_el(path):
    data = np.loadtxt(path, dtype=np.int64, ndmin=2)
    if data.size == 0:
        return np.zeros(0, np.int32), np.zeros(0, np.int32)
    return data[:, 0].astype(np.int32), data[:, 1].astype(np.int32)


def write_el(path, offsets, nbrs):
    offsets = np.asarray(offsets, np.int64)
    src = np.repeat(np.arange(len(offsets) - 1, dtype=np.int64), np.diff(offsets))
    with open(path, "w") as f:
        for u, v in zip(src.tolist(), np.asarray(nbrs).tolist()):
            f.write(f"{u} {v}\n")


def _inverse(n, offsets, nbrs):
    src = np.repeat(np.arange(n, dtype=np.int32), np.diff(offsets))
    order = np.lexsort((src, nbrs))                 # by destination, then source (ascending lists)
    in_off = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(nbrs, minlength=n), out=in_off[1:])
    return in_off, src[order].astype(np.int32)


def write_sg(path, offsets, nbrs, directed=False):
    offsets = np.ascontiguousarray(offsets, np.int64)
    nbrs = np.ascontiguousarray(nbrs, np.int32)
    n = len(offsets) - 1
    with open(path, "wb") as f:
        f.write(np.array([directed], np.bool_).tobytes())
        f.write(np.array([len(nbrs), n], np.int64).tobytes())
        f.write(offsets.tobytes())
        f.write(nbrs.tobytes())
        if directed:
            in_off, in_nbr = _inverse(n, offsets, nbrs)
            f.write(in_off.tobytes())
            f.write(in_nbr.tobytes())


def read_sg(path):
    """Returns (directed, offsets int64[n+1], nbrs int32[slots]); the inverse of a directed graph is not needed here."""
    with open(path, "rb") as f:
        directed = bool(np.frombuffer(f.read(1), np.bool_)[0])
        slots, n = np.frombuffer(f.read(16), np.int64)
        offsets = np.frombuffer(f.read(8 * (int(n) + 1)), np.int64).copy()
        nbrs = np.frombuffer(f.read(4 * int(slots)), np.int32).copy()
    if len(offsets) != n + 1 or len(nbrs) != slots or offsets[-1] != slots:
        raise ValueError(f"{path}: truncated or inconsistent .sg file")
    return directed, offsets, nbrs


def load_graph(path, symmetrize=True):
    """Builder::MakeGraph for a file (gapbs/builder.h:1642-1660): device graph from .sg (as is) or .el (built on GPU)."""
    from .capi import Graph
    if path.endswith(".sg"):
        directed, off, nbr = read_sg(path)
        return Graph.from_csr(off, nbr, directed)
    if path.endswith(".el"):
        src, dst = read_el(path)
        return Graph.from_edgelist(src, dst, symmetrize)
    raise ValueError(f"Unrecognized suffix: {path}")         # reader.h:243-245
