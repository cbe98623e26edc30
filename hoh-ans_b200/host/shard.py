"""Image-batch sharding across the GPUs of one box.

Images are independent units, so the batch is partitioned by contiguous ranges of image indices, one
range per rank; there is no data-path collective.  The only cross-rank steps are a barrier around the
timed region, the MAX of the per-rank device times and the SUM of the per-rank byte counts, done with
torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
import os


def world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process = 1 GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_range(n_items, rank, world_size):
    """Strong-scaling split: rank g takes items [g*N/G, (g+1)*N/G) (choh.cpp images are independent)."""
    lo = (n_items * rank) // world_size
    hi = (n_items * (rank + 1)) // world_size
    return lo, hi


def weak_first_seed(images_per_rank, rank):
    """Weak-scaling assignment used by bench.py: every rank codes its own `images_per_rank` images;
    image k of rank r is the global image r*images_per_rank + k, whose generator seed is 1 + that."""
    return 1 + rank * images_per_rank


def reduce_max(value, world_size, device=None):
    if world_size == 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value, world_size, device=None):
    if world_size == 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_sizes(sizes, world_size):
    """Final gather of the per-rank stream-size tables (the only cross-GPU exchange of the path):
    returns the list of all ranks' tables on every rank."""
    if world_size == 1:
        return [list(sizes)]
    import torch.distributed as dist
    out = [None] * world_size
    dist.all_gather_object(out, list(sizes))
    return out
