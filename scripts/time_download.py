"""Where the time of the result download goes (one rank, 403^3 fp32 grid -> float64 host array):
DMA alone, widening alone (library pool, by thread count; torch's cast copy), the pipelined path by chunk count,
and engine.to_host_f64 itself (pooled copy-on-write segment).  usage: python scripts/time_download.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from giwaxsim_b200 import _lib, engine

V = 403
n = V ** 3
dev = torch.device("cuda", 0)
t = torch.rand(n, device=dev)
stage = torch.empty(n, dtype=torch.float32, pin_memory=True)
dst = torch.empty(n, dtype=torch.float64)
dst.zero_()


def timed(f, reps=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps


def dma():
    stage.copy_(t, non_blocking=True); torch.cuda.synchronize()


print("cpus %d; DMA alone %.2f ms" % (os.cpu_count(), timed(dma)))
for thr in (2, 4, 8, 12, 16, 24, 32):
    print("widen alone, library pool %2d threads: %.2f ms" % (
        thr, timed(lambda: _lib.call("gx_host_widen_f32_f64", stage.data_ptr(), dst.data_ptr(), n, thr))))
torch.set_num_threads(min(16, os.cpu_count()))
print("widen alone, torch cast copy %d threads: %.2f ms" % (torch.get_num_threads(), timed(lambda: dst.copy_(stage))))


def pipelined(chunks, thr):
    step = -(-n // chunks)
    marks = []
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        stage[lo:hi].copy_(t[lo:hi], non_blocking=True)
        ev = torch.cuda.Event(); ev.record(); marks.append((lo, hi, ev))
    for lo, hi, ev in marks:
        ev.synchronize()
        _lib.call("gx_host_widen_f32_f64", stage.data_ptr() + 4 * lo, dst.data_ptr() + 8 * lo, hi - lo, thr)


for chunks in (4, 8, 16, 32):
    for thr in (8, 16):
        print("pipelined %2d chunks, %2d threads: %.2f ms" % (chunks, thr, timed(lambda: pipelined(chunks, thr))))
g = t.view(V, V, V)
keep = []
def full():
    keep.append(engine.to_host_f64(g))
    if len(keep) > 1:
        keep.pop(0)
print("engine.to_host_f64: %.2f ms" % timed(full, reps=6))
