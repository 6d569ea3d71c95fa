"""Timing of the reference's own call pattern (per-image similarity_map x2, cat, L1, backward:
realesrganssl_model.py:378-420) on the benchmark workload, next to the fused SelfSimilarityLoss call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssl_b200
from ssl_b200 import synth, functional as F_

dev = torch.device("cuda:0")
sr, gt, mask = synth.make_case(16, 256, 256, seed=1, density=0.114)
sr, gt, mask = sr.to(dev), gt.to(dev), mask.to(dev)


def reference_loop(path):
    x = sr.clone().requires_grad_(True)
    a, b = [], []
    for i in range(x.shape[0]):
        m = mask[i:i + 1]
        if m.sum() == 0:
            continue
        el = ssl_b200.build_edge_list(m)
        a.append(ssl_b200.ssg_rows(x[i:i + 1], el, 25, 9, 0.004, True, path=path).unsqueeze(0))
        b.append(ssl_b200.ssg_rows(gt[i:i + 1], el, 25, 9, 0.004, True, path=path).unsqueeze(0))
    loss = 1e3 * torch.nn.functional.l1_loss(torch.cat(a, dim=1), torch.cat(b, dim=1))
    loss.backward()
    return loss


def fused():
    x = sr.clone().requires_grad_(True)
    loss = ssl_b200.ssl(x, gt, mask, 25, 9, 0.004, True, loss_weight=1e3)
    loss.backward()
    return loss


def timeit(fn, iters=5):
    fn(); fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        out = fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters, float(out.detach())


for name, fn in (("per-image loop, point kernels", lambda: reference_loop("point")),
                 ("per-image loop, plane kernels", lambda: reference_loop("plane")),
                 ("fused SelfSimilarityLoss (plane)", fused)):
    ms, val = timeit(fn)
    print(f"{name:36s} {ms:8.2f} ms/step   loss {val:.6e}")
