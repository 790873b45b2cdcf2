"""CPU suite for the N>1 path: world_size-2 gloo run of the shard-and-reduce plumbing (gms_b200/dist.py).
The per-rank partial counts come from the CPU oracle here (there is no GPU); on the GPU box the same
plumbing is fed by gmsb_tc_total_ex(part_index, part_count) — see test_partition_sums_to_total."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from gms_b200 import dist as gd
    from oracle import binding
    r, w, _ = gd.init(backend="gloo")
    assert (r, w) == (rank, world)
    orc = binding.oracle()
    g = orc.generate(10)
    # shard the undirected edges i % world == rank, exactly like the device schedule is sharded
    sec, edges, partial = g.tc_total_sample(world, rank)
    total3, nedges = gd.allreduce_counts([partial, edges])
    tmax = gd.allreduce_max(float(rank + 1))
    gd.barrier()
    q.put((rank, partial, total3, nedges, tmax, gd.part_size(g.slots // 2, rank, world) == edges))


def test_two_rank_shard_and_allreduce(golden):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    rec = golden["generated"]["kronecker-10"]
    assert all(r[2] == 3 * rec["tc"] for r in res)            # Σ_{u<v}|N(u)∩N(v)| = 3·TC on every rank
    assert all(r[3] == rec["slots"] // 2 for r in res)
    assert res[0][1] + res[1][1] == 3 * rec["tc"] and res[0][1] != res[1][1]
    assert all(r[4] == 2.0 for r in res) and all(r[5] for r in res)


def test_part_size_covers_everything():
    from gms_b200.dist import part_size
    for total in (0, 1, 7, 1000):
        for parts in (1, 2, 3, 8):
            assert sum(part_size(total, i, parts) for i in range(parts)) == total


def _upload_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch
    from gms_b200 import dist as gd
    gd.init(backend="gloo")
    ok = True
    for n, slots in ((7, 23), (1, 0), (1000, 12345)):          # ragged: lengths not divisible by the world size
        off = torch.arange(n + 1, dtype=torch.int64) * 3
        nbr = (torch.arange(max(slots, 0), dtype=torch.int32) * 7) % 1001
        up = gd.ShardedCsrUpload(off, nbr, "cpu")
        for _ in range(2):                                       # buffers are reused
            o, b = up.upload()
            ok &= bool(torch.equal(o, off) and torch.equal(b, nbr))
        ok &= up.h2d_bytes <= 8 * ((n + 1 + world - 1) // world) + 4 * ((slots + world - 1) // world)
    gd.barrier()
    q.put((rank, ok))


def test_sharded_csr_upload_replicates_the_host_arrays():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_upload_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


# ---- the exchange of the sharded build (gms_b200.dist.ShardedOrientedBuild) over gloo ----------------------------------
class NumpyShard:
    """CPU stand-in with the interface of gms_b200.capi.Shard that restates the gmsb_shard_* protocol in numpy
    (gms_b200/csrc/graph_build.cu: shard_begin / shard_export / shard_finish): vertex cuts of equal slot counts, rows of
    the own range oriented by (degree, id) rank and packed in original-id order, rows of all pieces moved into rank order
    with  start(u) = scan[u] - scan[cut[r]] + r * stride."""

    def __init__(self, off, nbr, part, parts, offsets_dev_ptr=None):
        self.off, self.nbr, self.part, self.parts = off, nbr, part, parts
        n = self.n = len(off) - 1
        if offsets_dev_ptr is not None:                 # the all-gathered copy of the offsets: must equal the host's
            import ctypes
            got = np.ctypeslib.as_array((ctypes.c_int64 * (n + 1)).from_address(offsets_dev_ptr))
            assert np.array_equal(got, off)
        else:
            assert parts == 1
        last = int(off[n])
        self.cut = [0]
        for i in range(1, parts):
            self.cut.append(max(self.cut[-1], int(np.searchsorted(off[:n], last // parts * i, "left"))))
        self.cut.append(n)
        deg = np.diff(off)
        order = np.lexsort((np.arange(n), deg))                    # (degree asc, id asc)
        self.rank = np.empty(n, np.int64)
        self.rank[order] = np.arange(n)
        u0, u1 = self.cut[part], self.cut[part + 1]
        self.rows = []
        for u in range(u0, u1):
            r = self.rank[nbr[off[u]:off[u + 1]]]
            self.rows.append(np.sort(r[r > self.rank[u]]).astype(np.int32))
        self.piece_len = int(sum(len(r) for r in self.rows))

    @staticmethod
    def _view(ptr, count):
        import ctypes
        return np.ctypeslib.as_array((ctypes.c_int32 * max(count, 1)).from_address(ptr))[:count]

    def export(self, piece_ptr, dplus_all_ptr):
        piece = self._view(piece_ptr, self.piece_len)
        if self.rows:
            piece[:] = np.concatenate(self.rows) if self.piece_len else piece
        dplus = self._view(dplus_all_ptr, self.n)
        u0 = self.cut[self.part]
        dplus[u0:u0 + len(self.rows)] = [len(r) for r in self.rows]

    def finish(self, pieces_ptr, stride, dplus_all_ptr):
        n = self.n
        dplus = self._view(dplus_all_ptr, n).astype(np.int64)
        pieces = self._view(pieces_ptr, stride * self.parts)
        scan = np.concatenate([[0], np.cumsum(dplus)])
        doff = np.zeros(n + 1, np.int64)
        doff[self.rank + 1] = dplus
        doff = np.cumsum(doff)
        dnbr = np.empty(int(doff[n]), np.int32)
        for r in range(self.parts):
            for u in range(self.cut[r], self.cut[r + 1]):
                start = int(scan[u] - scan[self.cut[r]] + r * stride)
                dnbr[doff[self.rank[u]]:doff[self.rank[u]] + dplus[u]] = pieces[start:start + dplus[u]]
        return doff, dnbr


def _sharded_build_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch
    from gms_b200 import dist as gd
    from oracle import binding
    gd.init(backend="gloo")
    orc = binding.oracle()
    ok = True
    for scale in (6, 9):
        g = orc.generate(scale)
        off, nbr = g.csr()
        want_off, want_nbr = g.induce_directed(g.degree_order(True)).csr()
        build = gd.ShardedOrientedBuild(torch.from_numpy(off.astype(np.int64)), torch.from_numpy(nbr.astype(np.int32)),
                                        "cpu", shard_factory=NumpyShard)
        for _ in range(2):                                           # buffers are reused
            doff, dnbr = build.build()
            m = int(doff[-1])
            ok &= bool(np.array_equal(doff[:len(want_off)], want_off) and np.array_equal(dnbr[:m], want_nbr[:m]))
            ok &= bool(np.all(doff[len(want_off) - 1:] == m))        # vertices past the last endpoint have empty rows
    gd.barrier()
    q.put((rank, ok))


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_oriented_build_exchange(world):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_build_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
