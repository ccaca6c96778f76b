"""Stage B alone on a synthetic voxel grid (config 5 geometry): filtered vs exact kernel."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from giwaxsim_b200 import engine, synth
from giwaxsim_b200.tools import comparison
P = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n_or = int(sys.argv[2]) if len(sys.argv) > 2 else 360
generic = len(sys.argv) > 3
cfg = synth.config5()
dev = engine.resolve_device()
V = 403
q = np.linspace(-2.01, 2.01, V)
iq = torch.rand(V, V, V, device=dev)
gx, gy, gz, _, _ = comparison.detector_base_device(P, 2.0, cfg["angle_init_vals"], cfg["angle_init_axs"], dev)
psis = np.linspace(0, 89.75, n_or)
phis = np.array([7.3]) if generic else np.array([0.0])
R, w = engine.orientation_tables(engine.grid_corners(gx, gy, gz), psis, np.ones(n_or) / n_or, phis, np.ones(1), [0.0], np.ones(1))
det = engine.DetectorEngine(iq, q, q, q)
for exact in (True, False):
    img = torch.zeros(P * P, dtype=torch.float64, device=dev)
    det.accumulate(gx, gy, gz, R, w, image=img, exact_only=exact)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): det.accumulate(gx, gy, gz, R, w, image=img, exact_only=exact, count_slow=not exact)
    e1.record(); torch.cuda.synchronize()
    print("exact" if exact else "filtered", "%.3f ms per %d orientations" % (e0.elapsed_time(e1) / 3, n_or), "slow fraction", det.last_slow_fraction)
