"""world_size-2 gloo test of the N>1 host logic: round-robin sharding of phi
slices / orientations and the all-reduce that combines per-rank partial grids
(giwaxsim_b200/parallel.py).  The partial grids come from the CPU oracle here;
on GPUs the same code path carries device tensors over NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from giwaxsim_b200 import parallel
from oracle import giwaxs_oracle as ox


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, coords, f, r, q, max_q, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert parallel.rank_world() == (rank, world)
        setup = ox.stage_a_setup(coords, f, r, q, max_q)
        mine = parallel.shard(setup["phis"], rank, world)
        q3 = (setup["q_num"],) * 3
        vsum, vcnt = np.zeros(q3), np.zeros(q3)
        for phi in mine:
            ox.run_slice(vsum, vcnt, coords, setup, r, phi, True, 3)
        t_sum = torch.from_numpy(vsum)
        t_cnt = torch.from_numpy(vcnt.astype(np.int32))
        parallel.all_reduce_sum([t_sum, None, t_cnt])
        # replicated host array -> "device" copy, each rank contributing only its slice
        old = parallel.SHARDED_UPLOAD_MIN_BYTES
        parallel.SHARDED_UPLOAD_MIN_BYTES = 64
        try:
            for arr in (coords, np.arange(1001, dtype=np.int32), np.array(["C", "Si", "O"] * 37)):
                src = arr if arr.dtype.kind != "U" else np.ascontiguousarray(arr).view(np.uint32).view(np.int32)
                got = parallel.upload_replicated(src, torch.device("cpu"))
                assert got.shape == src.shape and np.array_equal(got.numpy(), src)
        finally:
            parallel.SHARDED_UPLOAD_MIN_BYTES = old
        # replicated result -> one float64 host copy per node, mapped copy-on-write by every rank
        os.environ["LOCAL_WORLD_SIZE"] = str(world)
        grid = torch.arange(5 * 7 * 3, dtype=torch.float32).reshape(5, 7, 3) * 0.5

        def widen(dev_slice, view):
            view[:] = dev_slice.numpy().astype(np.float64)

        shared = parallel.shared_result_f64(grid, widen, min_bytes=0)
        assert shared is not None and shared.dtype == np.float64 and shared.shape == (5, 7, 3)
        assert np.array_equal(shared, grid.numpy().astype(np.float64))
        shared[rank, 0, 0] = -1.0 - rank                       # private pages: the other rank must not see this
        dist.barrier()
        assert shared[1 - rank, 0, 0] == grid[1 - rank, 0, 0].item()
        assert not [f for f in os.listdir("/dev/shm") if f.startswith("giwaxs_b200_%d_" % port)]
        # pool: a segment is reused only when every rank dropped its array; live arrays never change
        second = parallel.shared_result_f64(grid * 3, widen, min_bytes=0)       # `shared` alive -> new segment
        segs = parallel._pool[grid.numel() * 8]
        assert len(segs) == 2 and np.array_equal(second, 3 * grid.numpy().astype(np.float64))
        keep = shared[2:4]                                                  # a view keeps its segment busy
        del shared
        third = parallel.shared_result_f64(grid * 5, widen, min_bytes=0)
        assert len(segs) == 3 and keep[0, 1, 1] == grid[2, 1, 1].item()
        del keep, second
        import gc
        gc.collect()
        fourth = parallel.shared_result_f64(grid * 7, widen, min_bytes=0)       # segments 0 and 1 are free again
        assert len(segs) == 3 and np.array_equal(fourth, 7 * grid.numpy().astype(np.float64))
        assert np.array_equal(third, 5 * grid.numpy().astype(np.float64))
        if rank == 1:
            hold = parallel.shared_result_f64(grid, widen, min_bytes=0)         # (kept alive on rank 1 only)
        else:
            parallel.shared_result_f64(grid, widen, min_bytes=0)
        fifth = parallel.shared_result_f64(grid * 9, widen, min_bytes=0)        # all three segments busy somewhere
        assert fifth is None
        # another result size gets its own class and does not evict the first one
        small = torch.arange(11, dtype=torch.float32)
        other = parallel.shared_result_f64(small, widen, min_bytes=0)
        assert np.array_equal(other, small.numpy().astype(np.float64)) and len(parallel._pool) == 2
        assert len(parallel._pool[grid.numel() * 8]) == 3
        os.environ["LOCAL_WORLD_SIZE"] = "1"                   # ranks "on different nodes": caller converts itself
        assert parallel.shared_result_f64(grid, widen, min_bytes=0) is None
        if rank == 0:
            ret["sum"], ret["cnt"], ret["n"] = t_sum.numpy().copy(), t_cnt.numpy().copy(), len(mine)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_reduce_equal_serial():
    rng = np.random.default_rng(2)
    coords = rng.random((200, 3)) * [12.0, 9.0, 10.0]
    f = np.full(200, 6.0049 + 0.0023j)
    r, q, max_q = 0.3, 0.2, 1.0
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, _free_port(), coords, f, r, q, max_q, ret), nprocs=2, join=True)
        got_sum, got_cnt, n0 = ret["sum"], ret["cnt"], ret["n"]
    _, _, _, _, vsum, vcnt, setup = ox.voxelgridmaker(coords, f, r, q, max_q, True, 3)
    assert n0 == len(setup["phis"][0::2])
    assert np.array_equal(got_cnt, vcnt.astype(np.int32))               # integer counts: exact
    assert np.abs(got_sum - vsum).max() <= 1e-12 * vsum.max()           # float sums: association only


def test_shard_is_a_partition():
    items = np.arange(23)
    parts = [parallel.shard(items, r, 4) for r in range(4)]
    assert sorted(np.concatenate(parts).tolist()) == items.tolist()
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
