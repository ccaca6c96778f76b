"""Multi-GPU plumbing: one process per GPU, torch.distributed for the exchange.

The path shards without any data-path collective: phi slices (stage A) and
detector orientations (stage B) are independent units that only meet in a sum
(reference: shared-memory `+=` at tools/voxelgrids.py:502-503 and
tools/detector.py:298).  Each rank therefore accumulates its round-robin share
into private grids and one all-reduce (NCCL over NVLink/NVSwitch on GPUs, gloo
in the CPU tests) combines them.  Integer counts stay exact under summation.
"""
import numpy as np
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


SHARDED_UPLOAD_MIN_BYTES = 8 << 20


def upload_replicated(array, device, group=None):
    """Device copy of a host array that every rank holds identically (the script is replicated
    under torchrun, as the reference's single process would run it).  With more than one rank each
    rank pushes only its 1/world slice over PCIe and an all-gather over NVLink completes the copy,
    so the host->device time of the atom table does not stay constant as GPUs are added.
    Returns a tensor of the array's dtype and shape on `device`."""
    a = np.ascontiguousarray(array)
    flat = torch.from_numpy(a.reshape(-1).view(np.uint8))
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    n = flat.numel()
    if world == 1 or n < SHARDED_UPLOAD_MIN_BYTES:
        out = flat.to(device, non_blocking=True)
    else:
        rank = dist.get_rank(group)
        per = ((n + world - 1) // world + 15) // 16 * 16
        full = torch.empty(per * world, dtype=torch.uint8, device=device)
        lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
        mine = full[rank * per:(rank + 1) * per]
        if hi > lo:
            mine[:hi - lo].copy_(flat[lo:hi], non_blocking=True)
        dist.all_gather_into_tensor(full, mine.clone() if dist.get_backend(group) == "gloo" else mine, group=group)
        out = full[:n]
    torch_dtype = torch.from_numpy(np.empty(0, dtype=a.dtype)).dtype
    return out.view(torch_dtype).view(a.shape)
