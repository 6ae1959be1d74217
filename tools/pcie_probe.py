import time, torch
n = 1 << 28  # 2 GiB of float64
h = torch.empty(n, dtype=torch.float64, pin_memory=True)
d = torch.empty(n, dtype=torch.float64, device="cuda")
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(name, "GB/s", n * 8 / dt / 1e9)
t0 = time.perf_counter(); x = torch.empty(1 << 30, dtype=torch.float64, pin_memory=True); print("pin 8GiB s", time.perf_counter() - t0)
