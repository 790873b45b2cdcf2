"""TEST INFRASTRUCTURE ONLY — ctypes bindings for the CPU checkers.

Two libraries export the same function set under different prefixes:

* ``liboracle.so``       (prefix ``orc_``)    — our restatement, oracle/oracle.cpp
* ``_ref/libgmsref.so``  (prefix ``gmsref_``) — the unmodified reference behind oracle/ref_shim.cpp

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
this module; nothing under gms_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libgmsref.so")

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")

METRICS = {"jaccard": 0, "overlap": 1, "adamic_adar": 2, "resource": 3, "comm_neigh": 4, "total_neigh": 5,
           "pref_att": 6}


def build(force=False):
    """Compile liboracle.so (always possible) and _ref/libgmsref.so (only where /root/reference exists)."""
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
            os.path.join(HERE, "oracle.cpp")):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/gms"):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


class Graph:
    def __init__(self, lib, handle):
        self.lib, self.h = lib, handle

    def __del__(self):
        if getattr(self, "h", None):
            self.lib._f("free")(self.h)
            self.h = None

    @property
    def n(self):
        return self.lib._f("num_nodes")(self.h)

    @property
    def slots(self):
        return self.lib._f("num_slots")(self.h)

    @property
    def directed(self):
        return bool(self.lib._f("directed")(self.h))

    def csr(self):
        off = np.zeros(self.n + 1, np.int64)
        nbr = np.zeros(max(self.slots, 1), np.int32)
        self.lib._f("export_csr")(self.h, off, nbr)
        return off, nbr[:off[-1]]

    # --- graph transforms
    def worth_relabelling(self):
        return bool(self.lib._f("worth_relabelling")(self.h))

    def relabel_by_degree(self):
        return Graph(self.lib, self.lib._f("relabel_by_degree")(self.h))

    def induce_directed(self, ranking):
        ranking = np.ascontiguousarray(ranking, np.int32)
        return Graph(self.lib, self.lib._f("induce_directed")(self.h, ranking))

    def degree_order(self, rank_format=False):
        out = np.zeros(self.n, np.int32)
        self.lib._f("degree_order")(self.h, int(rank_format), out)
        return out

    # --- kernels
    def tc_total(self, par=True):
        return int(self.lib._f("tc_total")(self.h, int(par)))

    def tc_total_timed(self, par=True):
        out = C.c_uint64(0)
        sec = self.lib._f("tc_total_timed")(self.h, int(par), C.byref(out))
        return sec, int(out.value)

    def tc_total_sample(self, stride, phase=0):
        edges, total = C.c_int64(0), C.c_uint64(0)
        sec = self.lib._f("tc_total_sample")(self.h, stride, phase, C.byref(edges), C.byref(total))
        return sec, int(edges.value), int(total.value)

    def tc_vertex2(self, variant=1):
        out = np.zeros(self.n, np.int64)
        self.lib._f("tc_vertex2")(self.h, variant, out)
        return out

    def tc_verify_total(self):
        return int(self.lib._f("tc_verify_total")(self.h))

    def kclique(self, k, mode=2):
        return int(self.lib._f("kclique")(self.h, k, mode))

    def kclique_timed(self, k, mode=2):
        out = C.c_uint64(0)
        sec = self.lib._f("kclique_timed")(self.h, k, mode, C.byref(out))
        return sec, int(out.value)

    def clique_count_set_based(self, k):
        return int(self.lib._f("clique_count_set_based")(self.h, k))

    def vertex_similarity(self, metric, a, b):
        return float(self.lib._f("vertex_similarity")(self.h, METRICS[metric], a, b))

    def pair_similarity(self, metric, a, b):
        a = np.ascontiguousarray(a, np.int32)
        b = np.ascontiguousarray(b, np.int32)
        out = np.zeros(len(a), np.float64)
        self.lib._f("pair_similarity")(self.h, METRICS[metric], len(a), a, b, out)
        return out

    def edge_similarity(self, metric):
        m = self.slots // 2
        out = np.zeros(max(m, 1), np.float64)
        k = self.lib._f("edge_similarity")(self.h, METRICS[metric], out)
        return out[:k]


class CpuLib:
    """One of the two CPU libraries, selected by prefix."""

    def __init__(self, path, prefix):
        self.path, self.prefix = path, prefix
        self.dll = C.CDLL(path)
        sig = {
            "set_threads": (None, [C.c_int]),
            "max_threads": (C.c_int, []),
            "generate_el": (None, [C.c_int, C.c_int, C.c_int, _i32p, _i32p]),
            "generate": (C.c_void_p, [C.c_int, C.c_int, C.c_int]),
            "from_el": (C.c_void_p, [C.c_int64, _i32p, _i32p, C.c_int]),
            "from_csr": (C.c_void_p, [C.c_int64, _i64p, _i32p, C.c_int]),
            "free": (None, [C.c_void_p]),
            "num_nodes": (C.c_int64, [C.c_void_p]),
            "num_slots": (C.c_int64, [C.c_void_p]),
            "directed": (C.c_int, [C.c_void_p]),
            "export_csr": (None, [C.c_void_p, _i64p, _i32p]),
            "worth_relabelling": (C.c_int, [C.c_void_p]),
            "relabel_by_degree": (C.c_void_p, [C.c_void_p]),
            "intersect_count": (C.c_uint64, [_i32p, C.c_int64, _i32p, C.c_int64]),
            "intersect": (C.c_int64, [_i32p, C.c_int64, _i32p, C.c_int64, _i32p]),
            "union": (C.c_int64, [_i32p, C.c_int64, _i32p, C.c_int64, _i32p]),
            "union_count": (C.c_uint64, [_i32p, C.c_int64, _i32p, C.c_int64]),
            "difference": (C.c_int64, [_i32p, C.c_int64, _i32p, C.c_int64, _i32p]),
            "contains": (C.c_int, [_i32p, C.c_int64, C.c_int32]),
            "tc_total": (C.c_uint64, [C.c_void_p, C.c_int]),
            "tc_total_timed": (C.c_double, [C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]),
            "tc_total_sample": (C.c_double, [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_int64),
                                             C.POINTER(C.c_uint64)]),
            "tc_vertex2": (None, [C.c_void_p, C.c_int, _i64p]),
            "tc_verify_total": (C.c_uint64, [C.c_void_p]),
            "degree_order": (None, [C.c_void_p, C.c_int, _i32p]),
            "induce_directed": (C.c_void_p, [C.c_void_p, _i32p]),
            "kclique": (C.c_uint64, [C.c_void_p, C.c_int, C.c_int]),
            "kclique_timed": (C.c_double, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_uint64)]),
            "clique_count_set_based": (C.c_uint64, [C.c_void_p, C.c_int]),
            "vertex_similarity": (C.c_double, [C.c_void_p, C.c_int, C.c_int32, C.c_int32]),
            "pair_similarity": (None, [C.c_void_p, C.c_int, C.c_int64, _i32p, _i32p, _f64p]),
            "edge_similarity": (C.c_int64, [C.c_void_p, C.c_int, _f64p]),
        }
        if prefix == "orc_":
            sig.update({
                "rmat_el": (None, [C.c_int, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_int, _i32p, _i32p]),
                "degeneracy_rank": (None, [C.c_void_p, _i32p]),
                "check_degeneracy_rank": (C.c_int64, [C.c_void_p, _i32p]),
                "core_number_of_rank": (C.c_int64, [C.c_void_p, _i32p]),
                "adg_order": (None, [C.c_void_p, C.c_double, C.c_int, _i32p, _i32p]),
                "adg_order_ex": (None, [C.c_void_p, C.c_double, C.c_int, C.c_int, _i32p, _i32p]),
                "clique_counts_pivot": (None, [C.c_void_p, C.c_int, np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")]),
                "tc_bytes": (None, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                    C.POINTER(C.c_int64)]),
                "kclique_bytes": (C.c_uint64, [C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]),
            })
        else:
            sig.update({"degeneracy_danisch_heap": (None, [C.c_void_p, _i32p]),
                        "adg_order": (None, [C.c_void_p, C.c_double, C.c_int, _i32p]),
                        "adg_order_ex": (None, [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, _i32p]),
                        "load_file": (C.c_void_p, [C.c_char_p, C.c_int]),
                        "write_file": (None, [C.c_void_p, C.c_char_p, C.c_int])})
        self._fn = {}
        for name, (res, args) in sig.items():
            f = getattr(self.dll, prefix + name)
            f.restype, f.argtypes = res, args
            self._fn[name] = f

    def _f(self, name):
        return self._fn[name]

    # --- threads
    def set_threads(self, t):
        self._f("set_threads")(t)

    def max_threads(self):
        return self._f("max_threads")()

    # --- graph sources
    def generate_el(self, scale, degree=16, uniform=False):
        m = (1 << scale) * degree
        src, dst = np.zeros(m, np.int32), np.zeros(m, np.int32)
        self._f("generate_el")(scale, degree, int(uniform), src, dst)
        return src, dst

    def rmat_el(self, scale, m, a, b, c, permute=True):
        src, dst = np.zeros(m, np.int32), np.zeros(m, np.int32)
        self._f("rmat_el")(scale, m, a, b, c, int(permute), src, dst)
        return src, dst

    def generate(self, scale, degree=16, uniform=False):
        return Graph(self, self._f("generate")(scale, degree, int(uniform)))

    def from_el(self, src, dst, symmetrize=True):
        src = np.ascontiguousarray(src, np.int32)
        dst = np.ascontiguousarray(dst, np.int32)
        return Graph(self, self._f("from_el")(len(src), src, dst, int(symmetrize)))

    def from_csr(self, off, nbr, directed=False):
        off = np.ascontiguousarray(off, np.int64)
        nbr = np.ascontiguousarray(nbr, np.int32)
        if len(nbr) == 0:
            nbr = np.zeros(1, np.int32)
        return Graph(self, self._f("from_csr")(len(off) - 1, off, nbr, int(directed)))

    # --- set algebra
    @staticmethod
    def _arr(x):
        x = np.ascontiguousarray(x, np.int32)
        return (x if len(x) else np.zeros(1, np.int32)), len(x)

    def intersect_count(self, a, b):
        (a, na), (b, nb) = self._arr(a), self._arr(b)
        return int(self._f("intersect_count")(a, na, b, nb))

    def _binary(self, name, a, b):
        (a, na), (b, nb) = self._arr(a), self._arr(b)
        out = np.zeros(na + nb + 1, np.int32)
        k = self._f(name)(a, na, b, nb, out)
        return out[:k].copy()

    def intersect(self, a, b):
        return self._binary("intersect", a, b)

    def union(self, a, b):
        return self._binary("union", a, b)

    def difference(self, a, b):
        return self._binary("difference", a, b)

    def union_count(self, a, b):
        (a, na), (b, nb) = self._arr(a), self._arr(b)
        return int(self._f("union_count")(a, na, b, nb))

    def contains(self, a, x):
        a, na = self._arr(a)
        return bool(self._f("contains")(a, na, x))

    # --- reference-only: graph files through the reference's own reader / writer
    def load_file(self, path, symmetrize=True):
        return Graph(self, self._f("load_file")(path.encode(), int(symmetrize)))

    def write_file(self, g, path, serialized=True):
        self._f("write_file")(g.h, path.encode(), int(serialized))

    # --- oracle-only
    def degeneracy_rank(self, g):
        out = np.zeros(g.n, np.int32)
        name = "degeneracy_rank" if self.prefix == "orc_" else "degeneracy_danisch_heap"
        self._f(name)(g.h, out)
        return out

    def check_degeneracy_rank(self, g, rank):
        return int(self._f("check_degeneracy_rank")(g.h, np.ascontiguousarray(rank, np.int32)))

    def adg_order(self, g, eps=1.0, rank_format=False):
        """Approximate degeneracy order (averageDegree boundary). The oracle also returns the round of each vertex."""
        out = np.zeros(max(g.n, 1), np.int32)
        if self.prefix == "orc_":
            rounds = np.zeros(max(g.n, 1), np.int32)
            self._f("adg_order")(g.h, float(eps), int(rank_format), out, rounds)
            return out[:g.n], rounds[:g.n]
        self._f("adg_order")(g.h, float(eps), int(rank_format), out)
        return out[:g.n]

    def adg_order_ex(self, g, eps=1.0, rank_format=False, boundary="average", pull=False):
        """ADG with a boundary function ("average" | "min"); the reference also takes pull=True (the Set form)."""
        kind = {"average": 0, "min": 1}[boundary]
        out = np.zeros(max(g.n, 1), np.int32)
        if self.prefix == "orc_":
            rounds = np.zeros(max(g.n, 1), np.int32)
            self._f("adg_order_ex")(g.h, float(eps), int(rank_format), kind, out, rounds)
            return out[:g.n], rounds[:g.n]
        self._f("adg_order_ex")(g.h, float(eps), int(rank_format), kind, int(pull), out)
        return out[:g.n]

    def clique_counts_pivot(self, dag, kmax):
        """counts[k] for k = 0..kmax on an oriented DAG, by pivoting (independent cross-check, oracle only)."""
        out = np.zeros(kmax + 1, np.uint64)
        self._f("clique_counts_pivot")(dag.h, kmax, out)
        return [int(x) for x in out]

    def core_number_of_rank(self, g, rank):
        return int(self._f("core_number_of_rank")(g.h, np.ascontiguousarray(rank, np.int32)))

    def kclique_bytes(self, dag, k):
        """(B_k, count): algorithmic bytes of the set-algebra clique recursion on the DAG and the count."""
        cnt = C.c_uint64(0)
        b = self._f("kclique_bytes")(dag.h, k, C.byref(cnt))
        return int(b), int(cnt.value)

    def tc_bytes(self, g):
        bt, br, mx = C.c_uint64(0), C.c_uint64(0), C.c_int64(0)
        self._f("tc_bytes")(g.h, C.byref(bt), C.byref(br), C.byref(mx))
        return int(bt.value), int(br.value), int(mx.value)


_cache = {}


def oracle():
    """Our CPU restatement (always available after build())."""
    if "orc" not in _cache:
        if not os.path.exists(ORACLE_SO):
            build()
        _cache["orc"] = CpuLib(ORACLE_SO, "orc_")
    return _cache["orc"]


def reference():
    """The unmodified reference, or None when the prebuilt _ref/libgmsref.so is absent."""
    if "ref" not in _cache:
        _cache["ref"] = CpuLib(REF_SO, "gmsref_") if os.path.exists(REF_SO) else None
    return _cache["ref"]
