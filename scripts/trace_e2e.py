"""Where does the host time of voxelgridmaker_fitting / detectormaker_fitting go? (config 5)"""
import sys, time, cProfile, pstats
import numpy as np, torch
sys.path.insert(0, ".")
from giwaxsim_b200 import synth
from giwaxsim_b200.tools import comparison, utilities
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
out = a(); b(*out); torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable(); t0 = time.perf_counter(); out = a(); torch.cuda.synchronize(); t1 = time.perf_counter(); pr.disable()
print("stage A wall %.3f s" % (t1 - t0)); pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
pr = cProfile.Profile(); pr.enable(); t0 = time.perf_counter(); b(*out); torch.cuda.synchronize(); t1 = time.perf_counter(); pr.disable()
print("stage B wall %.3f s" % (t1 - t0)); pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
