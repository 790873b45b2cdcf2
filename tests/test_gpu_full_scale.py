"""GPU parity (-m gpu) at BASELINE.json's full size (configs[1]: Kronecker scale 24, edge factor 16).

The reference would need about an hour of CPU for the exact total at this size, so the checks are the
size-independent ones: four independent kernel families must agree (bitmap, merge path, galloping, and the per-edge
support kernels through Σ vertex_count2 = 6·TC), the answer is invariant under relabelling and additive over the
multi-GPU partition, the builder reproduces the survey's n / m / max degree / max d+ exactly, and a random sample of
per-edge intersection counts is compared with the CPU oracle on the same (full-size) graph."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kron24(gms):
    s, d = gms.generate_rmat(24)
    g = gms.Graph.from_edgelist(s, d, True)
    del s, d
    return g


def test_builder_matches_survey_at_scale_24(kron24):
    g = kron24
    assert g.n == 16777212 and g.slots // 2 == 260376709
    off, _ = g.export_csr()
    assert int(np.diff(off).max()) == 406979
    _, st = g.tc_total_ex(reuse_plan=True)
    assert st["max_dplus"] == 1762 and st["oriented_edges"] == 260376709
    assert st["algorithmic_bytes"] == 1076700748208                  # B_TC = 1.0767 TB (exact)


def test_kernel_families_agree_at_scale_24(kron24):
    g = kron24
    total, _ = g.tc_total_ex(reuse_plan=True)
    assert total == 10283205554
    assert g.tc_total_ex(variant="gallop")[0] == total
    assert g.tc_total_ex(variant="merge")[0] == total
    parts = [g.tc_total_ex(part_index=p, part_count=8, reuse_plan=True)[0] for p in range(8)]
    assert sum(parts) == total and max(parts) < 1.2 * min(parts)     # additive and balanced
    v2 = g.tc_vertex2()
    assert int(v2.sum()) == 6 * total and (v2 % 2 == 0).all()


def test_sampled_edges_against_oracle_at_scale_24(kron24, orc):
    g = kron24
    off, nbr = g.export_csr()
    o = orc.from_csr(off, nbr, False)
    rng = np.random.default_rng(24)
    slots = rng.integers(0, len(nbr), 20000)
    a = (np.searchsorted(off, slots, side="right") - 1).astype(np.int32)
    b = nbr[slots].astype(np.int32)
    want = o.pair_similarity("comm_neigh", a, b).astype(np.uint64)
    assert (g.intersect_count_batch(a, b) == want).all()
    for m in ("jaccard", "overlap", "total_neigh", "pref_att", "resource"):
        assert g.pair_similarity(m, a, b).tobytes() == o.pair_similarity(m, a, b).tobytes(), m
    # the hubs themselves: the 64 highest-degree vertices against each other (lists of 10^5 elements, many merge tiles)
    hubs = np.argsort(np.diff(off))[-64:].astype(np.int32)
    ha, hb = np.meshgrid(hubs, hubs)
    ha, hb = ha.ravel().astype(np.int32), hb.ravel().astype(np.int32)
    assert (g.intersect_count_batch(ha, hb) == o.pair_similarity("comm_neigh", ha, hb).astype(np.uint64)).all()


def test_skewed_rmat_scale_20_families_agree(gms):
    """R-MAT a = 0.65 at scale 20 (max d+ near a thousand, much heavier hubs than the kronecker constants give): the bitmap schedule, the
    all-gallop schedule, the partition sums and the per-vertex counts agree."""
    s, d = gms.generate_rmat(20, a=0.65, b=0.15, c=0.15)
    g = gms.Graph.from_edgelist(s, d, True)
    auto, st = g.tc_total_ex(reuse_plan=True)
    assert st["edges_bitmap"] > 0 and st["max_dplus"] > 500
    assert g.tc_total_ex(variant="gallop")[0] == auto
    assert sum(g.tc_total_ex(part_index=p, part_count=3, reuse_plan=True)[0] for p in range(3)) == auto
    assert int(g.tc_vertex2().sum()) == 6 * auto
