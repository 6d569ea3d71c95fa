"""Dev check of the plane forward against the point kernels and the oracle (run under gpurun)."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ssl_b200
from ssl_b200 import _lib, synth, functional as F_
from oracle import ssl_oracle as oracle

dev = torch.device("cuda:0")
vp = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())


def plane_rows(img, img2, el, n, ks, kw):
    b, c, h, w = img.shape
    rows = torch.empty(n, ks * ks, device=dev)
    rows2 = torch.empty_like(rows) if img2 is not None else None
    nb = int(_lib.load().ssl_b200_plane_rows_workspace_bytes(b, h, w, ks, kw, n))
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.call("ssl_b200_plane_rows_forward", vp(img), vp(img2), _lib.dtype_code(img.dtype), b, c, h, w, vp(el.edges),
              vp(el.counts), n, ks, kw, vp(rows), vp(rows2), vp(ws), nb, st)
    return rows, rows2


for (B, H, W, ks, kw, rho) in [(2, 20, 24, 7, 3, 0.2), (2, 64, 64, 11, 5, 0.1), (1, 48, 56, 25, 9, 0.05),
                               (2, 100, 130, 25, 9, 0.114)]:
    sr, gt, mask = synth.make_case(B, H, W, seed=3, density=rho)
    el = ssl_b200.build_edge_list(mask.to(dev))
    n = el.count()
    rows, rows2 = plane_rows(sr.to(dev), gt.to(dev), el, n, ks, kw)
    torch.cuda.synchronize()
    pr, pr2 = F_._rows_forward(sr.to(dev), gt.to(dev), el, n, ks, kw, 1.0, 0.0, 0)
    ref = []
    for i in range(B):
        pos = oracle.edge_positions(mask[i, 0].numpy())
        ref.append(oracle.raw_distance(sr[i].numpy().astype(np.float64), pos, ks, kw))
    ref = np.concatenate(ref)
    got = rows.cpu().numpy()
    err = np.abs(got - ref) / (np.abs(ref) + 1e-6)
    errp = np.abs(pr.cpu().numpy() - ref) / (np.abs(ref) + 1e-6)
    print(f"B{B} {H}x{W} ks{ks} kw{kw}: n={n} plane max rel err {err.max():.2e} (point {errp.max():.2e}); "
          f"gt vs point {float((rows2 - pr2).abs().max()):.2e}")
    if err.max() > 1e-4:
        bad = np.argwhere(err > 1e-4)
        print("  bad entries", len(bad), "first", bad[:5], "delta idx ->", [(int(d) // ks - ks // 2, int(d) % ks - ks // 2) for _, d in bad[:5]])

# timing at config 2
sr, gt, mask = synth.make_case(16, 256, 256, seed=1, density=0.114)
sr, gt, mask = sr.to(dev), gt.to(dev), mask.to(dev)
el = ssl_b200.build_edge_list(mask)
n = el.count()
for _ in range(2):
    plane_rows(sr, gt, el, n, 25, 9)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    plane_rows(sr, gt, el, n, 25, 9)
e.record(); torch.cuda.synchronize()
print(f"config2 plane rows fwd x2 (incl. lists, eout, transpose): {s.elapsed_time(e) / 5:.3f} ms, n={n}")
