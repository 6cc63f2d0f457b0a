"""Host-side partitioning of independent work across GPUs.

Batched systems and multi-RHS blocks shard by contiguous index ranges, one process per GPU; there is no
collective on the data path -- torch.distributed is only used to agree on the elapsed time (max over
ranks) and to gather small results.  A single triangular solve never spans GPUs ("replicas only").
"""
import os


def shard_range(total, rank, world):
    """Contiguous, balanced [lo, hi) slice of `total` units for `rank` of `world` (GPU g of G owns
    [g*total/G, (g+1)*total/G))."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return (total * rank) // world, (total * (rank + 1)) // world


def dist_env():
    """(rank, local_rank, world_size) from the torchrun environment (1 process = 1 GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_process_group(backend=None):
    """Join the job's process group if WORLD_SIZE > 1 (nccl on GPUs, gloo on CPU); returns (rank, local, world)."""
    rank, local, world = dist_env()
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (identity for a single process)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
