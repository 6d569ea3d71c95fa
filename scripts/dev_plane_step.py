"""Dev check of the whole plane step (fwd + row loss + bwd) against the oracle and the point path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ssl_b200
from ssl_b200 import synth
from oracle import ssl_oracle as oracle

dev = torch.device("cuda:0")
cases = [(2, 20, 24, 7, 3, 0.2, 0.0), (2, 64, 64, 11, 5, 0.1, 0.0), (1, 48, 56, 25, 9, 0.05, 0.0),
         (2, 100, 130, 25, 9, 0.114, 0.5), (1, 16, 40, 25, 9, 0.3, 0.0)]
if os.environ.get("QUICK"):
    cases = cases[:2]
for (B, H, W, ks, kw, rho, wkl) in cases:
    sr, gt, mask = synth.make_case(B, H, W, seed=3, density=rho)
    res = {}
    for path in ("point", "plane"):
        x = sr.to(dev).requires_grad_(True)
        total, l1, kl = ssl_b200.ssl(x, gt.to(dev), mask.to(dev), ks, kw, 0.004, True, loss_weight=1.0, kl_weight=wkl,
                                     return_parts=True, path=path)
        total.backward()
        torch.cuda.synchronize()
        res[path] = (float(l1), float(kl), x.grad.cpu().numpy())
    l1o, klo, go, n = oracle.loss_and_grad(sr.numpy().astype(np.float64), gt.numpy().astype(np.float64), mask.numpy(),
                                           ks, kw, 0.004, True, kl_weight=wkl)
    gmax = np.abs(go).max()
    for path in ("point", "plane"):
        l1, kl, g = res[path]
        print(f"B{B} {H}x{W} ks{ks} kw{kw} n={n} {path:5s}: l1 rel {abs(l1 - l1o) / l1o:.2e} "
              f"kl rel {abs(kl - klo) / max(klo, 1e-30):.2e} grad err {np.abs(g - go).max() / gmax:.2e}")
    if np.abs(res['plane'][2] - go).max() / gmax > 1e-4:
        d = np.abs(res['plane'][2] - go) / gmax
        idx = np.argwhere(d > 1e-4)
        print("   bad px:", len(idx), idx[:8].tolist())

if not os.environ.get("QUICK"):
    sr, gt, mask = synth.make_case(16, 256, 256, seed=1, density=0.114)
    sr, gt, mask = sr.to(dev), gt.to(dev), mask.to(dev)
    for path in ("point", "plane"):
        x = sr.clone().requires_grad_(True)
        def step():
            x.grad = None
            ssl_b200.ssl(x, gt, mask, 25, 9, 0.004, True, path=path).backward()
        step(); step(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5): step()
        e.record(); torch.cuda.synchronize()
        print(f"config2 {path}: {s.elapsed_time(e) / 5:.3f} ms/step")
