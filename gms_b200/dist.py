"""Multi-GPU plumbing: one process per GPU, oriented edges partitioned, one all-reduce of the counts.

Counting shards with no data-path collective: every rank takes share `rank` of `world` of the edges
(gmsb_tc_options.part_index / part_count: those whose closing vertex it owns; it builds the schedule of that share
only) and the uint64 partial counts are summed by a single all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests).

Getting a HOST CSR onto N devices is where collectives pay: `ShardedOrientedBuild` has every rank upload and orient one
vertex range and all-gathers the finished rows of the oriented representation (what the triangle kernels read);
`ShardedCsrUpload` replicates the symmetric arrays themselves (rank r uploads slice r of N, one all-gather per array) for
the operators that need them.  Either way the host is read once, not N times.
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment; returns (rank, world, local_rank)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local


def allreduce_counts(values, device=None):
    """Sum a short vector of non-negative integer counts (< 2^63) over all ranks; returns Python ints."""
    vals = [int(v) for v in values]
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return vals
    assert all(0 <= v < (1 << 63) for v in vals)
    t = torch.tensor(vals, dtype=torch.int64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(x) for x in t.tolist()]


def allreduce_max(value, device=None):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64,
                     device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def part_size(total, index, parts):
    """Number of schedule entries i in [0,total) with i % parts == index (mirrors tc.cu:part_size)."""
    return (total - index + parts - 1) // parts if total > index else 0


def tc_total_sharded(graph, **opts):
    """Triangle count with the schedule sharded over the ranks of the default process group."""
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    part, stats = graph.tc_total_ex(part_index=rank, part_count=world, **opts)
    total, = allreduce_counts([part])
    return total, part, stats


class ShardedOrientedBuild:
    """Host CSR -> oriented device graph with the work sharded over the ranks (gmsb_shard_*): rank r uploads 1/world of
    the offsets (all-gathered over NVLink: the ranking needs every degree) and the neighbour slots of vertex range r of
    `world`, orients those rows, the finished rows are all-gathered and the d+ values all-reduced, and every rank moves
    the rows into rank order.  Over the host links go 8(n+1) + 4*slots bytes in total, whatever the number of ranks.
    The graph `build()` returns answers the triangle entry points (its symmetric lists are incomplete by construction).
    Buffers are allocated once and reused by every `build()`."""

    def __init__(self, offsets_host, nbrs_host, device, shard_factory=None):
        self.rank, self.world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
        self.device = torch.device(device)
        self.shard_factory = shard_factory          # tests drive the exchange on CPU tensors with a stand-in
        self.off_t = offsets_host if isinstance(offsets_host, torch.Tensor) else torch.from_numpy(offsets_host)
        self.off = self.off_t.numpy()
        self.nbr = nbrs_host.numpy() if isinstance(nbrs_host, torch.Tensor) else nbrs_host
        self.n = len(self.off) - 1
        self.dplus = torch.zeros(max(self.n, 1), dtype=torch.int32, device=self.device)
        self.pieces = None
        self.off_upload = _ShardedArray(self.off_t, self.device, self.rank, self.world) if self.world > 1 else None

    @property
    def h2d_bytes(self):
        """Bytes this rank copies host->device per build (its share of the offsets + its vertex range's slots; the
        latter from the same cuts the library computes)."""
        if self.world == 1:
            return 8 * (self.n + 1) + 4 * int(self.off[self.n])
        last = int(self.off[self.n])
        cut = [0] + [int(np.searchsorted(self.off[:self.n], last // self.world * i, "left")) for i in range(1, self.world)]
        cut = list(np.maximum.accumulate(cut)) + [self.n]
        own = int(self.off[cut[self.rank + 1]] - self.off[cut[self.rank]])
        return self.off_upload.h2d_bytes + 4 * own

    def build(self):
        if self.shard_factory is None:
            from . import capi
            self.shard_factory = capi.Shard
        off_dev = self.off_upload.upload().data_ptr() if self.off_upload is not None else None
        shard = self.shard_factory(self.off, self.nbr, self.rank, self.world, offsets_dev_ptr=off_dev)
        stride = max(int(allreduce_max(shard.piece_len, device=self.device)), 1)
        if self.pieces is None or self.pieces.numel() < stride * self.world:
            self.pieces = torch.empty(stride * self.world, dtype=torch.int32, device=self.device)
        pieces = self.pieces.narrow(0, 0, stride * self.world)
        mine = pieces.narrow(0, self.rank * stride, stride)
        self.dplus.zero_()
        shard.export(mine.data_ptr(), self.dplus.data_ptr())
        if self.world > 1:
            dist.all_gather_into_tensor(pieces, mine)
            dist.all_reduce(self.dplus, op=dist.ReduceOp.SUM)
        return shard.finish(pieces.data_ptr(), stride, self.dplus.data_ptr())


class _ShardedArray:
    """One host array replicated on every rank's device: rank r copies slice r of `world` host->device, one all-gather
    fills the rest.  The device buffer is padded to a multiple of `world` and reused by every `upload()`."""

    def __init__(self, host, device, rank, world):
        self.host, self.rank, self.world = host, rank, world
        self.len = host.numel()
        self.per = _slice_len(self.len, world)
        self.full = torch.empty(self.per * world, dtype=host.dtype, device=device)

    def span(self):
        lo = min(self.rank * self.per, self.len)
        return lo, min(self.per, self.len - lo)

    @property
    def h2d_bytes(self):
        return self.span()[1] * self.host.element_size()

    def upload(self):
        lo, cnt = self.span()
        mine = self.full.narrow(0, self.rank * self.per, self.per)
        if cnt:
            mine.narrow(0, 0, cnt).copy_(self.host.narrow(0, lo, cnt), non_blocking=True)
        if self.world > 1:
            dist.all_gather_into_tensor(self.full, mine)
        return self.full.narrow(0, 0, self.len)


def _slice_len(total, world):
    return (total + world - 1) // world


class ShardedCsrUpload:
    """Replicates a host CSR (pinned int64 offsets[n+1], int32 nbrs[slots]) on every rank's device: rank r copies slice
    r of `world` host->device, one all-gather per array fills the rest (NVLink on GPUs; gloo + CPU tensors in the CPU
    test).  Buffers are allocated once and reused by every `upload()`; arrays are padded to a multiple of `world`.
    (For operators that need the symmetric lists on every device; the triangle path uses ShardedOrientedBuild.)"""

    def __init__(self, offsets_host, nbrs_host, device):
        self.rank, self.world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
        self.arrays = [_ShardedArray(h, torch.device(device), self.rank, self.world) for h in (offsets_host, nbrs_host)]

    @property
    def h2d_bytes(self):
        """Bytes this rank copies host->device per upload."""
        return sum(a.h2d_bytes for a in self.arrays)

    def upload(self):
        """Returns (offsets_dev, nbrs_dev) views of the replicated arrays (valid until the next upload)."""
        return tuple(a.upload() for a in self.arrays)
