"""Multi-GPU plumbing for the tensor-stream job: one process per GPU.

The matrix tensor shards naturally (every element is an independent product against the same
fixed operand): rank g takes the contiguous index range
[g*N/G, (g+1)*N/G).  The only exchange on the path is the one-off broadcast of the fixed
operand (NCCL over NVLink on GPUs; gloo in the CPU tests); results never cross ranks.
The reference has no multi-GPU code at all (only count_available_gpus, src/cudamatrix.cc:14-18).
"""
import os


def shard_range(count, rank, world):
    """Contiguous, balanced, order-preserving partition of range(count)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    lo = (count * rank) // world
    hi = (count * (rank + 1)) // world
    return lo, hi


def dist_env():
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def broadcast_fixed(fixed, src=0, group=None):
    """Broadcast the fixed operand (a torch tensor, device or CPU) from `src` to every rank, in
    place.  One collective per job; with the nccl backend this is the NVLink/NVSwitch broadcast."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(fixed, src=src, group=group)
    return fixed


def gather_counts(local_count, group=None):
    """All ranks learn every rank's processed count (for the whole-job throughput figure)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [local_count]
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.tensor([local_count], dtype=torch.int64, device=dev)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return [int(x.item()) for x in out]
