"""Per-kernel CUDA-event timing of the staged (unfused) pipeline on a synthetic slab.
usage: python scripts/time_stages.py [n_atoms] [grid] [n_phi]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from giwaxsim_b200 import engine, synth
from giwaxsim_b200.tools import utilities

utilities.set_f1f2_provider(synth.fixed_f1f2)
n_atoms = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
n_phi = int(sys.argv[3]) if len(sys.argv) > 3 else 8
cfg = synth.config5()
q = 0.01
r = 2 * np.pi / (q * (N - 0.5))
box = (560.0 * N / 4096, 250.0 * N / 4096, 560.0 * N / 4096)
coords, el = synth.random_slab(n_atoms, box)
t0 = time.time()
codes, uniq = engine.encode_values(el)
print("encode %.3fs" % (time.time() - t0))
table = [complex(*synth.fixed_f1f2(str(e))) + utilities.ATOMIC_NUMBER[str(e)] for e in uniq]
dev = engine.resolve_device()
torch.cuda.synchronize(); t0 = time.time()
atoms = engine.AtomSet(coords, r, N, dev, species=codes, table=table)
torch.cuda.synchronize(); print("upload+sort %.3fs" % (time.time() - t0))
Ncheck, q_num, q_axis, phis = engine.stage_a_geometry(atoms.bounds, r, q, 2.0)
assert Ncheck == N
sum_f = np.sum(np.bincount(codes, minlength=len(table)) * np.asarray(table))
avg = sum_f / np.prod(atoms.bounds) * r ** 3
eng = engine.SliceEngine(None, r, q_axis, N, avg, atoms.bounds[0], atoms.bounds[1], True, 25, atoms=atoms)
sel = phis[:: max(1, len(phis) // n_phi)][:n_phi]
grid = torch.empty(len(sel) * N * N * 2, dtype=torch.float32, device=dev)
work = torch.empty_like(grid); iq2d = torch.empty(len(sel) * N * N, dtype=torch.float32, device=dev)

def timed(name, fn, reps=3):
    fn(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps): fn()
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    print("%-12s %8.3f ms / batch of %d  -> %8.1f us/slice" % (name, ms, len(sel), 1e3 * ms / len(sel)))
    return ms

t = {}
def prep():
    t["t"] = eng.prepare(sel)
timed("prepare", prep)
timed("project", lambda: eng.project(t["t"], grid))
timed("fft2", lambda: eng.fft(grid, work, iq2d, len(sel)))
timed("bin", lambda: eng.bin(t["t"], iq2d))
print("atoms", n_atoms, "N", N, "q_num", q_num, "slices", len(sel), "bbox", t["t"]["bbox"][:4].tolist())
work = torch.empty(len(sel) * N * eng.KC * 2, dtype=torch.float32, device=dev)
timed("fused", lambda: eng.fused(t["t"], work))
print("KC", eng.KC, "rows kept", eng.row_lo, eng.row_hi)
