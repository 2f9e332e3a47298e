import torch, time
x = torch.empty(84_700_000 // 4, device="cuda")
h = torch.empty(84_700_000 // 4).pin_memory()
h2 = torch.empty(84_700_000 // 4).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    a = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - a) / n
d = t(lambda: h.copy_(x, non_blocking=True))
print(f"D2H 84.7 MB one stream: {84.7e6 / d / 1e9:.1f} GB/s")
half = x.numel() // 2
def two():
    with torch.cuda.stream(s1): h[:half].copy_(x[:half], non_blocking=True)
    with torch.cuda.stream(s2): h[half:].copy_(x[half:], non_blocking=True)
d = t(two)
print(f"D2H two streams halves: {84.7e6 / d / 1e9:.1f} GB/s")
def both():
    with torch.cuda.stream(s1): h.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): x[:2_100_000].copy_(h2[:2_100_000], non_blocking=True)
d = t(both)
print(f"D2H 84.7 MB with concurrent H2D 8.4 MB: {84.7e6 / d / 1e9:.1f} GB/s")
