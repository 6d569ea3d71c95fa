"""Quick timing probe of the point kernels at config 2 (dev tool, not the benchmark)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssl_b200
from ssl_b200 import synth, functional as F_

dev = torch.device("cuda:0")
B = int(os.environ.get("B", 16))
sr, gt, mask = synth.make_case(B, 256, 256, seed=1, density=float(os.environ.get("RHO", 0.114)))
sr, gt, mask = sr.to(dev), gt.to(dev), mask.to(dev)
el = ssl_b200.build_edge_list(mask)
n = el.count()
print("edge px", n)

def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters

t = timeit(lambda: ssl_b200.build_edge_list(mask))
print(f"edge list      {t:8.3f} ms")
t = timeit(lambda: F_._rows_forward(sr, gt, el, n, 25, 9, 0.004, 1e-10, 2))
print(f"rows fwd x2    {t:8.3f} ms  {n / t / 1e3:8.2f} M edge-px/s")
rows, rows2 = F_._rows_forward(sr, gt, el, n, 25, 9, 0.004, 1e-10, 2)
gq = torch.randn_like(rows)
t = timeit(lambda: F_._rows_backward(sr, el, n, 25, 9, gq))
print(f"rows bwd       {t:8.3f} ms  {n / t / 1e3:8.2f} M edge-px/s")
x = sr.clone().requires_grad_(True)
def step():
    x.grad = None
    ssl_b200.ssl(x, gt, mask, 25, 9, 0.004, True).backward()
t = timeit(step)
print(f"ssl fwd+bwd    {t:8.3f} ms  {n / t / 1e3:8.2f} M edge-px/s")
