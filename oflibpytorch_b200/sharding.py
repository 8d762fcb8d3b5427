"""Batch sharding of the hot path across the GPUs of one node.

Every hot-path operation is independent per batch item, so the batch (or frame) dimension is split into contiguous
shards, one per rank / GPU, and each rank runs the ordinary single-GPU calls on its shard: there is NO collective on
the data path.  The only cross-batch coupling in the reference is its zero-flow early-out, which looks at the WHOLE
batch passed to one call (``all(is_zero_flow(...))``, utils.py:497, flow_class.py:1046,1729,1738); a driver that
wants bit-identical behaviour to the unsharded call ANDs the per-shard flags first (`agree_all_zero`, one 4-byte
all-reduce, off the data path) and, if the global answer differs from the local one, avoids the early-out by
calling the kernels through `Flow` objects whose flags it overrides (`force_nonzero`).
"""
from typing import List, Sequence, Tuple

import torch


def shard_range(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [start, stop) of `n_items` for `rank` (frames [256g/G, 256(g+1)/G))."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("invalid rank %d for world size %d" % (rank, world_size))
    return (n_items * rank) // world_size, (n_items * (rank + 1)) // world_size


def shard_batch(rank: int, world_size: int, *tensors):
    """Slice every tensor along dim 0 to this rank's shard; tensors with batch size 1 (broadcast) and None pass
    through unchanged."""
    out = []
    n = max((t.shape[0] for t in tensors if t is not None), default=0)
    a, b = shard_range(n, world_size, rank)
    for t in tensors:
        out.append(t if (t is None or t.shape[0] == 1) else t[a:b])
    return out if len(out) != 1 else out[0]


def agree_all_zero(local_all_zero: bool, group=None) -> bool:
    """Logical AND of the per-shard "whole shard is zero flow" flags over all ranks (what the unsharded call's
    early-out would have seen).  Works with any initialised process group (nccl or gloo); without one it is the
    identity."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return bool(local_all_zero)
    backend = dist.get_backend(group)
    dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    t = torch.tensor([1 if local_all_zero else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(int(t.item()))


def force_nonzero(flow) -> None:
    """Mark a Flow as "not all zero" so that the hot-path methods skip their early-out even when this shard's
    vectors happen to be zero (used when `agree_all_zero` says some OTHER shard is not)."""
    stats = flow._get_stats()
    flow._stats = [s | 2 | 4 | 8 | 16 for s in stats]


def gather_shards(local: torch.Tensor, group=None) -> List[torch.Tensor]:
    """Optional epilogue (off the timed path): all-gather equally sized per-rank results."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [local]
    out = [torch.empty_like(local) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, local.contiguous(), group=group)
    return out
