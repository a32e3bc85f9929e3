"""Host <-> device copy bandwidth of this box with page-locked memory: the ceiling of the e2e (host buffers) figure."""
import time, torch
n = 52428800
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print("%s %d MB pinned: %.3f ms, %.1f GB/s" % (name, n >> 20, dt * 1e3, n / dt / 1e9))
# chunked like the pipeline: 12 copies of 1/12
c = n // 12
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    for i in range(12): h[i * c:(i + 1) * c].copy_(d[i * c:(i + 1) * c], non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print("D2H in 12 chunks: %.3f ms, %.1f GB/s" % (dt * 1e3, n / dt / 1e9))
