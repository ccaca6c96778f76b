"""SURVEY section 7, T9 on hardware: the result of N ranks (one process per GPU, NCCL over NVLink)
equals the result of one rank.  The reference accumulates every slice / orientation into ONE shared
grid (tools/voxelgrids.py:502-503, tools/detector.py:298); here each rank accumulates its round-robin
share and the partial grids are combined by the collective in giwaxsim_b200/parallel.py.

Integer counts must be identical bit for bit; float grids / images may differ only by summation
order (<= 1e-6 of the maximum).  Skipped on a box with a single GPU (`gpurun --gpus 2` runs it);
bench.py performs the same comparison on every multi-rank run (`check.multi_gpu_vs_1rank`).
"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

R_VOXEL, MAX_Q = 0.3, 1.5
DET_ARGS = (96, MAX_Q, (90.0, 90.0, 90.0), ("psi", "phi", "psi"))
PSIS, PHIS, THETAS = np.linspace(60, 90, 7), np.linspace(0, 170, 5), np.array([0.0, 2.0])


def _inputs():
    from giwaxsim_b200 import synth
    coords, elements = synth.random_slab(60_000, (120.0, 70.0, 110.0), seed=5)
    return coords, elements, synth.pow2_q_voxel(R_VOXEL, 512)


def _pipeline():
    from giwaxsim_b200 import synth
    from giwaxsim_b200.tools import comparison, utilities
    utilities.set_f1f2_provider(synth.fixed_f1f2)
    coords, elements, q = _inputs()
    iq, qx, qy, qz, eng = comparison.voxelgridmaker_fitting(coords, elements, R_VOXEL, q, MAX_Q, 12700.0,
                                                            fill_bkg=True, smooth=5, return_state=True)
    det, _, _ = comparison.detectormaker_fitting(iq, qx, qy, qz, *DET_ARGS, PSIS, None, PHIS, None, THETAS, None)
    return {"iq": np.array(iq), "det": np.array(det), "count2": eng.count2.cpu().numpy().copy(),
            "row_hist": eng.row_hist.cpu().numpy().copy(), "slices": int(eng.slices_done)}


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), LOCAL_WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        out = _pipeline()
        ret[rank] = out
    finally:
        from giwaxsim_b200 import parallel
        parallel.shutdown()
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_n_rank_result_equals_one_rank_result(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    one = _pipeline()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
        got = {k: ret[k] for k in range(world)}
    total = sum(got[k]["slices"] for k in range(world))
    assert total == one["slices"]                                   # the shards partition the rotation list
    for k in range(world):
        g = got[k]
        assert np.array_equal(g["count2"], one["count2"])             # integer counts: bit-exact
        assert np.array_equal(g["row_hist"], one["row_hist"])
        assert np.abs(g["iq"] - one["iq"]).max() <= 1e-6 * one["iq"].max()
        assert np.abs(g["det"] - one["det"]).max() <= 1e-6 * one["det"].max()
    # every rank returns the same arrays (replicated result)
    assert np.array_equal(got[0]["iq"], got[world - 1]["iq"])
    assert np.array_equal(got[0]["det"], got[world - 1]["det"])
