"""Host profile of voxelgridmaker_fitting / detectormaker_fitting under torchrun (config 5): rank 0 prints a
cProfile of one steady-state call and the wall time of call vs the barrier that follows it.
usage: torchrun --nproc-per-node N scripts/trace_e2e_multi.py"""
import cProfile, os, pstats, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
from giwaxsim_b200 import synth
from giwaxsim_b200.tools import comparison, utilities
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank()
utilities.set_f1f2_provider(synth.fixed_f1f2)
cfg = synth.config5()
coords, el = synth.random_slab(cfg["n_atoms"], cfg["box"])
phis = np.linspace(0, 179.9, 1800)
def a():
    return comparison.voxelgridmaker_fitting(coords, el, cfg["r_voxel_size"], cfg["q_voxel_size"], cfg["max_q"], 12700.0,
                                             fill_bkg=True, smooth=25, phis=phis)
def b(iq, qx, qy, qz):
    return comparison.detectormaker_fitting(iq, qx, qy, qz, 2048, 2.0, cfg["angle_init_vals"], cfg["angle_init_axs"],
                                            cfg["psis"], None, cfg["phis"], None, cfg["thetas"], None, mirror=True)
def barrier():
    dist.barrier(); torch.cuda.synchronize()
for _ in range(3):
    out = a(); barrier(); d = b(*out); barrier()
for it in range(4):
    barrier(); t0 = time.perf_counter(); out = a(); t1 = time.perf_counter(); barrier(); t2 = time.perf_counter()
    d = b(*out); t3 = time.perf_counter(); barrier(); t4 = time.perf_counter()
    print("rank %d call %d: A %.1f ms + barrier %.1f ms | B %.1f ms + barrier %.1f ms" % (
        rank, it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3)), flush=True)
barrier()
pr = cProfile.Profile()
if rank == 0: pr.enable()
out = a()
if rank == 0:
    pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
barrier()
pr = cProfile.Profile()
if rank == 0: pr.enable()
d = b(*out)
if rank == 0:
    pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
barrier()
dist.destroy_process_group()
