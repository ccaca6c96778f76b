"""Latency of ONE fit-loop evaluation at BASELINE configs[0] (PM6 slab, template config: N = 1048,
892 slices, 2880 orientations x 500^2) through the two drop-in drivers from host arrays - the call
`fit_slabsize.py:19-47` repeats per evaluation.  Slab from tests/golden/pm6.npz.
    python scripts/time_config1.py [reps] [profile]"""
import os, sys, time, cProfile, pstats
import numpy as np, torch
sys.path.insert(0, ".")
from giwaxsim_b200 import synth
from giwaxsim_b200.tools import comparison, utilities
utilities.set_f1f2_provider(synth.fixed_f1f2)
g = np.load("tests/golden/pm6.npz")
elements = np.array([str(n) for n in g["element_names"]])[g["element_codes"]]
coords = g["coords"]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5

def a():
    return comparison.voxelgridmaker_fitting(coords, elements, float(g["r"]), float(g["q"]), float(g["max_q"]), 12700.0,
                                             fill_bkg=True, smooth=int(g["smooth"]))
def b(iq, qx, qy, qz):
    return comparison.detectormaker_fitting(iq, qx, qy, qz, int(g["P"]), float(g["max_q"]), tuple(g["vals"]),
                                            tuple(str(x) for x in g["axs"]), g["psis"], None, g["phis"], None,
                                            g["thetas"], None, mirror=True)
for _ in range(2):
    out = a(); b(*out)
torch.cuda.synchronize()
ta, tb = [], []
for _ in range(reps):
    t0 = time.perf_counter(); out = a(); torch.cuda.synchronize(); t1 = time.perf_counter()
    b(*out); torch.cuda.synchronize(); t2 = time.perf_counter()
    ta.append(t1 - t0); tb.append(t2 - t1)
print("config 1: stage A %.1f ms (min %.1f), stage B %.1f ms (min %.1f) per evaluation; 892 slices, 2880 orientations"
      % (1e3 * np.median(ta), 1e3 * min(ta), 1e3 * np.median(tb), 1e3 * min(tb)))
if len(sys.argv) > 2:
    for name, fn in (("A", a), ("B", lambda: b(*out))):
        pr = cProfile.Profile(); pr.enable(); fn(); torch.cuda.synchronize(); pr.disable()
        print("---- stage", name); pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
