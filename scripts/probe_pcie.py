"""Dev helper: pinned host<->device copy rates of the box (what the e2e number can hide at best)."""
import torch, time
dev = torch.device("cuda:0")
for mb in (4, 6.3, 12.6, 29.4):
    n = int(mb * 1e6 / 4)
    h = torch.empty(n, dtype=torch.float32).pin_memory()
    d = torch.empty(n, dtype=torch.float32, device=dev)
    for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): fn()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print(f"{name} {mb:5.1f} MB: {ms:.3f} ms = {mb / ms:.1f} GB/s")
# both directions at once
h1 = torch.empty(int(12.6e6 / 4)).pin_memory(); d1 = torch.empty_like(h1, device=dev)
h2 = torch.empty(int(12.6e6 / 4)).pin_memory(); d2 = torch.empty_like(h2, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); print("H2D + D2H concurrently, 12.6 MB each:", (time.perf_counter() - t0) / 10 * 1e3, "ms")
