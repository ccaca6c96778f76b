"""Multi-GPU plumbing: one process per GPU, torch.distributed for the exchange.

The path shards without any data-path collective: phi slices (stage A) and
detector orientations (stage B) are independent units that only meet in a sum
(reference: shared-memory `+=` at tools/voxelgrids.py:502-503 and
tools/detector.py:298).  Each rank therefore accumulates its round-robin share
into private grids and one all-reduce (NCCL over NVLink/NVSwitch on GPUs, gloo
in the CPU tests) combines them.  Integer counts stay exact under summation.
"""
import torch
import torch.distributed as dist


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard(items, rank, world):
    """Round-robin share of `items` for `rank` (balances the phi-dependent
    number of kept columns across ranks)."""
    return items[rank::world]


def all_reduce_sum(tensors, group=None):
    """In-place sum over ranks of every tensor in the list (None entries skipped)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in tensors:
        if t is not None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
