"""Stage B alone on a synthetic voxel grid (config 5 geometry): the three detector kernels.

    python scripts/time_detector.py [P] [n_orient] [generic] [kernels,comma,separated]
"""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from giwaxsim_b200 import engine, synth
from giwaxsim_b200.tools import comparison
P = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n_or = int(sys.argv[2]) if len(sys.argv) > 2 else 360
generic = len(sys.argv) > 3 and sys.argv[3] == "generic"
kernels = sys.argv[4].split(",") if len(sys.argv) > 4 else ["exact", "filtered", "affine"]
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
ref = None
for k in kernels:
    img = torch.zeros(P * P, dtype=torch.float64, device=dev)
    det.accumulate(gx, gy, gz, R, w, image=img, kernel=k, count_slow=(k in ("filtered", "affine")))
    slow = det.last_slow_fraction
    if ref is None:
        ref = img.clone()
    err = float((img - ref).abs().max() / ref.abs().max())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3): det.accumulate(gx, gy, gz, R, w, image=img, kernel=k)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("%-8s %.3f ms per %d orientations (host prep included)  %.0f GB/s algorithmic  slow fraction %s  rel diff vs first %.2e  plan %s"
          % (k, ms, n_or, 4.0 * P * P * n_or / ms / 1e6, slow, err, det.last_plan if k == "affine" else ""))
