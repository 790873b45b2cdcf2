"""CPU suite for the N>1 path: world_size-2 gloo run of the shard-and-reduce plumbing (gms_b200/dist.py).
The per-rank partial counts come from the CPU oracle here (there is no GPU); on the GPU box the same
plumbing is fed by gmsb_tc_total_ex(part_index, part_count) — see test_partition_sums_to_total."""
import os
import socket
import sys

import numpy as np
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
