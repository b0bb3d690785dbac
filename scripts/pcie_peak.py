import torch, time
dev = torch.device("cuda", 0)
for mb in (1, 4, 22, 64, 256):
    n = mb << 20
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    for direction in ("d2h", "h2d"):
        for _ in range(3):
            (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps):
            (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        print("%s %4d MB  %.3f ms  %.1f GB/s" % (direction, mb, dt * 1e3, n / dt / 1e9))
# both directions at once
n = 64 << 20
d1 = torch.empty(n, dtype=torch.uint8, device=dev); h1 = torch.empty(n, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n, dtype=torch.uint8, device=dev); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    with torch.cuda.stream(s1): h1.copy_(d1, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print("duplex 64+64 MB %.3f ms  %.1f GB/s each" % (dt * 1e3, n / dt / 1e9))
