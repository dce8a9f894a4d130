"""Host-side plumbing for the multi-GPU configs.

C4 (batched independent worlds) shards naturally: rank r owns a contiguous block of worlds and steps
them with no data-path collective; torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used
only for the start/stop barrier and the max-over-ranks timing the benchmark reports."""
import os


def shard_range(n_items, rank, world_size):
    """Contiguous block partition: returns (first, count) of the items owned by `rank`."""
    base, rem = divmod(n_items, world_size)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend):
    import torch.distributed as dist
    rank, local_rank, world = dist_env()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def all_reduce_max(value, device="cpu"):
    """max over ranks of a python float (used for the timed region's duration)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_reduce_sum(value, device="cpu"):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
