import torch, time
n = 225*1024*1024//8*8
h = torch.empty(n, dtype=torch.float64, pin_memory=True); h.fill_(1.0)
d = torch.empty(n, dtype=torch.float64, device="cuda")
for name, f in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    f(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3): f()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t)/3
    print(name, "GB/s", n*8/dt/1e9)
