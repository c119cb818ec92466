import torch, time
x = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
y = torch.empty(32 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda"); dy = torch.empty(32 << 20, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(fn, n=10):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n
t = run(lambda: d.copy_(x, non_blocking=True)); print("H2D 256MB: %.2f ms %.1f GB/s" % (t * 1e3, 0.268435456 / t))
def both():
    with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): y.copy_(dy, non_blocking=True)
t = run(both); print("H2D 256MB + D2H 32MB concurrently: %.2f ms (H2D %.1f GB/s)" % (t * 1e3, 0.268435456 / t))
for mb in (1, 4, 16, 64):
    n = mb << 20
    t = run(lambda: d[:n].copy_(x[:n], non_blocking=True), 50); print("H2D %d MB: %.1f GB/s" % (mb, n / t / 1e9))
