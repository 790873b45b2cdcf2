"""The sharded build of the oriented representation (gmsb_shard_*, include/gmsb.h) against the ordinary build.

On one device the shards of all parts are built one after the other and their pieces placed where the all-gather of a
multi-process run would put them (the collectives themselves are torch.distributed's; tests/test_dist_cpu.py drives the
same exchange over gloo).  Every part's finished handle must then count exactly what gmsb_graph_from_csr counts:
totals for every variant, per-vertex counts, and the oriented CSR itself through the statistics."""
import numpy as np
import pytest

from conftest import random_graph_edges

pytestmark = pytest.mark.gpu


def build_sharded(gms, off, nbr, parts, device_offsets=False):
    import torch
    dev = torch.device("cuda", 0)
    n = len(off) - 1
    off_dev = torch.from_numpy(np.ascontiguousarray(off, np.int64)).to(dev) if device_offsets else None
    torch.cuda.synchronize()
    shards = [gms.capi.Shard(off, nbr, p, parts, offsets_dev_ptr=off_dev.data_ptr() if device_offsets else None)
              for p in range(parts)]
    stride = max(max(s.piece_len for s in shards), 1)
    pieces = torch.full((parts * stride,), -7, dtype=torch.int32, device=dev)        # padding must never be read
    dplus = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    for p, s in enumerate(shards):
        s.export(pieces.data_ptr() + 4 * p * stride, dplus.data_ptr())
    torch.cuda.synchronize()
    graphs = [s.finish(pieces.data_ptr(), stride, dplus.data_ptr()) for s in shards]
    return graphs, int(dplus.sum().item())


def check_against_plain(gms, src, dst, parts_list):
    g = gms.Graph.from_edgelist(src, dst, True)
    off, nbr = g.export_csr()
    want, st = g.tc_total_ex(reuse_plan=True)
    want_v2 = g.tc_vertex2() if g.n else None
    for i, parts in enumerate(parts_list):
        graphs, m = build_sharded(gms, off, nbr, parts, device_offsets=bool(i % 2))    # offsets from the host / the device
        assert m == st["oriented_edges"]
        for p, gs in enumerate(graphs):
            assert (gs.n, gs.slots, gs.directed) == (g.n, g.slots, False)
            for variant in ("auto", "merge", "gallop", "bitmap"):
                got, st2 = gs.tc_total_ex(variant=variant, reuse_plan=True)
                assert got == want, (parts, p, variant)
                assert st2["oriented_edges"] == st["oriented_edges"] and st2["max_dplus"] == st["max_dplus"]
                assert st2["algorithmic_bytes"] == st["algorithmic_bytes"]          # same lists behind every edge
        # shares taken from different parts' handles add up (what a multi-process run computes)
        assert sum(graphs[p].tc_total_ex(part_index=p, part_count=parts)[0] for p in range(parts)) == want
        if want_v2 is not None:
            assert np.array_equal(graphs[-1].tc_vertex2(), want_v2)
    return g


@pytest.mark.parametrize("scale", [8, 12, 16])
def test_sharded_build_counts_like_the_plain_build(gms, golden, scale):
    src, dst = gms.generate_rmat(scale)
    g = check_against_plain(gms, src, dst, (1, 2, 3, 8))
    assert g.tc_total() == golden["generated"][f"kronecker-{scale}"]["tc"]


def test_sharded_build_skewed_and_ragged(gms):
    # a=0.65 R-MAT: longer hub lists (mid / long sort queues, big-list path at scale 18), uneven ranges
    src, dst = gms.generate_rmat(18, a=0.65, b=0.15, c=0.15)
    check_against_plain(gms, src, dst, (2, 5))
    # more parts than vertices with edges, isolated tail vertices
    src, dst = random_graph_edges(5, 40, 200)
    check_against_plain(gms, np.append(src, 63).astype(np.int32), np.append(dst, 0).astype(np.int32), (1, 7, 64))
    # 70 vertices with 5000 + 69 neighbours each (lists of more than kBigList slots: the many-CTAs-per-list kernels) that
    # keep 0..69 of them after orientation (register and shared-memory sorters), next to 5000 vertices of degree 70
    left = np.arange(70, dtype=np.int32)
    right = np.arange(70, 5070, dtype=np.int32)
    bs, bd = np.repeat(left, len(right)), np.tile(right, len(left))
    cs, cd = np.triu_indices(70, 1)
    check_against_plain(gms, np.concatenate([bs, cs.astype(np.int32)]), np.concatenate([bd, cd.astype(np.int32)]), (1, 2, 4))


def test_sharded_handle_refuses_operators_that_need_the_symmetric_lists(gms):
    src, dst = gms.generate_rmat(10)
    g = gms.Graph.from_edgelist(src, dst, True)
    off, nbr = g.export_csr()
    (gs,), _ = build_sharded(gms, off, nbr, 1)
    assert gs.tc_total() == g.tc_total()
    for call in (lambda: gs.export_csr(), lambda: gs.kclique_count(4), lambda: gs.edge_similarity("jaccard"),
                 lambda: gs.degree_order(), lambda: gs.intersect_count_batch([0], [1])):
        with pytest.raises(gms.GmsbError):
            call()


def test_shard_arguments_are_checked(gms):
    off = np.array([0, 1, 2], np.int64)
    nbr = np.array([1, 0], np.int32)
    for part, parts in ((2, 2), (-1, 2), (0, 0), (0, 65)):
        with pytest.raises(gms.GmsbError):
            gms.capi.Shard(off, nbr, part, parts)
    with pytest.raises(gms.GmsbError):
        gms.capi.Shard(np.array([0, 5, 2], np.int64), nbr, 0, 1)            # offsets not monotone
    with pytest.raises(gms.GmsbError):
        gms.capi.Shard(off, np.array([1, 9], np.int32), 0, 1)               # id out of range
