"""Pinned host <-> device copy bandwidth of this box (what bounds the upload / download part of `e2e`)."""
import time, torch
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print("%s pinned 256 MiB: %.2f ms = %.1f GB/s" % (name, 1e3 * dt, n / dt / 1e9))
p = torch.empty(n, dtype=torch.uint8)
t0 = time.perf_counter(); d.copy_(p); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("H2D pageable 256 MiB: %.2f ms = %.1f GB/s" % (1e3 * dt, n / dt / 1e9))
s = torch.empty(n, dtype=torch.uint8)
for thr in (1, 4, 8, 16):
    torch.set_num_threads(thr)
    s.copy_(p); t0 = time.perf_counter(); s.copy_(p); dt = time.perf_counter() - t0
    print("host memcpy 256 MiB with %2d threads: %.2f ms = %.1f GB/s" % (thr, 1e3 * dt, n / dt / 1e9))
