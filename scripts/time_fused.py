"""Production stage-A path (windowed fused kernels) on the config-5 slab: one prepare + fused batches.
usage: python scripts/time_fused.py [n_atoms] [grid] [n_phi] [reps]   (ncu-friendly: few launches)"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from giwaxsim_b200 import engine, synth
from giwaxsim_b200.tools import comparison, utilities

utilities.set_f1f2_provider(synth.fixed_f1f2)
n_atoms = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
n_phi = int(sys.argv[3]) if len(sys.argv) > 3 else 64
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
q = 0.01
r = 2 * np.pi / (q * (N - 0.5))
box = (560.0 * N / 4096, 250.0 * N / 4096, 560.0 * N / 4096)
coords, el = synth.random_slab(n_atoms, box)
dev = engine.resolve_device()
codes, uniq, counts = engine.encode_elements_device(el, dev)
table = comparison.f_table(uniq, 12700.0)
atoms = engine.AtomSet(coords, r, N, dev, species=codes, table=table)
_, q_num, q_axis, phis = engine.stage_a_geometry(atoms.bounds, r, q, 2.0)
avg = np.sum(counts * np.asarray(table)) / np.prod(atoms.bounds) * r ** 3
window = engine.crop_range(q_axis, 2.0)
eng = engine.SliceEngine(None, r, q_axis, N, avg, atoms.bounds[0], atoms.bounds[1], True, 25, atoms=atoms,
                         window=window)
sel = phis[:: max(1, len(phis) // n_phi)][:n_phi]
eng.run(sel)                                    # warm-up (plans, candidates)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(reps):
    eng.run(sel)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("fused path: %.3f ms per %d slices -> %.1f us/slice; KC %d rows %d..%d window %s q_out %d"
      % (ms, len(sel), 1e3 * ms / len(sel), eng.KC, eng.row_lo, eng.row_hi, window, eng.q_out))
