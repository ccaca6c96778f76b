import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from giwaxsim_b200 import synth
from giwaxsim_b200.tools import comparison, utilities
utilities.set_f1f2_provider(synth.fixed_f1f2)
cfg = synth.config5()
coords, el = synth.random_slab(cfg["n_atoms"], cfg["box"])
phis = np.linspace(0, 179.9, 1800)
for i in range(4):
    t0 = time.perf_counter()
    out = comparison.voxelgridmaker_fitting(coords, el, cfg["r_voxel_size"], cfg["q_voxel_size"], cfg["max_q"], 12700.0, fill_bkg=True, smooth=25, phis=phis)
    t1 = time.perf_counter()
    d = comparison.detectormaker_fitting(*out, 2048, 2.0, cfg["angle_init_vals"], cfg["angle_init_axs"], cfg["psis"], None, cfg["phis"], None, cfg["thetas"], None, mirror=True)
    t2 = time.perf_counter()
    print("call %d: A %.1f ms, B %.1f ms" % (i, 1e3 * (t1 - t0), 1e3 * (t2 - t1)), flush=True)
    del out, d
