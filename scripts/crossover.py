"""Point vs plane kernels over mask density at the benchmark shape (picks SSL_B200_PATH_AUTO's threshold)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssl_b200
from ssl_b200 import synth

dev = torch.device("cuda:0")
for rho in (0.0025, 0.005, 0.01, 0.02, 0.04, 0.114, 0.344):
    sr, gt, mask = synth.make_case(16, 256, 256, seed=1, density=rho, )
    sr, gt, mask = sr.to(dev), gt.to(dev), mask.to(dev)
    out = []
    for path in ("point", "plane"):
        x = sr.clone().requires_grad_(True)
        def step():
            x.grad = None
            ssl_b200.ssl(x, gt, mask, 25, 9, 0.004, True, path=path).backward()
        step(); step(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(3): step()
        e.record(); torch.cuda.synchronize()
        out.append(s.elapsed_time(e) / 3)
    n = int(mask.sum())
    print(f"density {rho:6.4f}  edge px {n:7d}  point {out[0]:8.3f} ms  plane {out[1]:8.3f} ms  "
          f"plane {n / out[1] / 1e3:7.2f} M edge-px/s")
