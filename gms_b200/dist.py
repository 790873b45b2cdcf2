"""Multi-GPU plumbing: one process per GPU, oriented edges partitioned, CSR replicated, one all-reduce of the counts.

The path shards with no data-path collective: every rank builds the same graph, takes share
`rank` of `world` of the schedule (gmsb_tc_options.part_index / part_count) and the uint64 partial counts are
summed by a single all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import os

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
