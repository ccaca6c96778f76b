"""Where the time of one stage-B call goes (bench geometry), host wall clock with syncs."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from giwaxsim_b200 import engine, synth
from giwaxsim_b200.tools import comparison
cfg = synth.config5()
dev = engine.resolve_device()
P, n_or, V = 2048, int(sys.argv[1]) if len(sys.argv) > 1 else 360, 403
q = np.linspace(-2.01, 2.01, V)
iq = torch.rand(V, V, V, device=dev)
gx, gy, gz, _, _ = comparison.detector_base_device(P, 2.0, cfg["angle_init_vals"], cfg["angle_init_axs"], dev)
psis = np.linspace(0, 89.75, n_or)
R, w = engine.orientation_tables(engine.grid_corners(gx, gy, gz), psis, np.ones(n_or) / n_or, [0.0], np.ones(1), [0.0], np.ones(1))
image = torch.zeros(P * P, dtype=torch.float64, device=dev)
sync = torch.cuda.synchronize
def lap(name, fn, reps=20):
    fn(); sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    sync()
    print("%-28s %8.1f us" % (name, 1e6 * (time.perf_counter() - t0) / reps))
    return out
det = lap("DetectorEngine()", lambda: engine.DetectorEngine(iq, q, q, q, device=dev))
lap("image.zero_", lambda: image.zero_())
lap("_dev(R)", lambda: engine._dev(R, dev))
plan = lap("affine_plan (cached fit)", lambda: det.affine_plan(gx, gy, gz, R, w))
lap("_dev(records)", lambda: engine._dev(plan[1], dev))
lap("accumulate (all of it)", lambda: det.accumulate(gx, gy, gz, R, w, image=image))
lap("epilogue", lambda: engine.detector_epilogue(image, P, P, True, dev, finish=True))
def stage_b():
    image.zero_()
    d = engine.DetectorEngine(iq, q, q, q, device=dev)
    d.accumulate(gx, gy, gz, R, w, image=image)
    return engine.detector_epilogue(image, P, P, True, dev, finish=True)
lap("stage_b total", stage_b)
