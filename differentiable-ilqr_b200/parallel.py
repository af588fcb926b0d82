"""Multi-GPU plumbing: one process per GPU; independent MPC problems shard by
batch index with NO data-path collective; the imitation-learning loop all-reduces
one small gradient buffer per step (NCCL over NVLink on the GPU box, gloo in the
CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_batch, rank, world):
    """Contiguous, balanced [lo, hi) slice of the batch owned by `rank`."""
    base, rem = divmod(n_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(flat, group=None):
    """In-place SUM all-reduce of a small flat buffer (d theta, dq, dp, loss)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if flat.is_cuda and dist.get_backend(group) == "gloo":
            # gloo (CPU tests, two ranks sharing one GPU): stage through the host
            host = flat.cpu()
            dist.all_reduce(host, op=dist.ReduceOp.SUM, group=group)
            flat.copy_(host)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat
