import torch, time
n = 42 * 1024 * 1024 // 8 * 8
h = torch.empty(n // 8, dtype=torch.float64).pin_memory()
d = torch.empty(n // 8, dtype=torch.float64, device="cuda")
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): fn()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(name, "%.1f GB/s" % (20 * n / dt / 1e9))
