"""Cases for compute-sanitizer (memcheck / racecheck / synccheck) over the plane path and the drop-in entries:
    compute-sanitizer --tool memcheck python scripts/sanitize_cases.py"""
import ctypes, sys
import torch
sys.path.insert(0, ".")
import ssl_b200
from ssl_b200 import synth, _lib

dev = torch.device("cuda:0")
for (b, h, w, ks, kw, rho, dt, stride, kl) in ((2, 70, 66, 25, 9, 0.2, torch.float32, 0, 1.0),
                                              (1, 16, 40, 25, 9, 0.3, torch.bfloat16, 0, 0.0),
                                              (2, 40, 44, 11, 5, 0.15, torch.float32, 3, 1.0),
                                              (1, 30, 30, 7, 3, 1.0, torch.float32, 0, 0.0)):
    sr, gt, mask = synth.make_case(b, h, w, seed=5, density=rho)
    x = sr.to(dev).to(dt).requires_grad_(True)
    loss = ssl_b200.ssl(x, gt.to(dev), mask.to(dev), ks, kw, kl_weight=kl, mask_stride=stride, path="plane")
    loss.backward()
    el = ssl_b200.build_edge_list(mask.to(dev))
    y = sr.to(dev).requires_grad_(True)
    rows = ssl_b200.ssg_rows(y, el, ks, kw, path="plane")
    rows.sum().backward()
    torch.cuda.synchronize()
    print("ok", b, h, w, ks, kw, float(loss))
# drop-in entries (reference convention) on the plane kernels
sr, _, mask = synth.make_case(1, 64, 72, seed=2, density=0.2)
P = 12
img_pad = torch.nn.functional.pad(sr[0], (P, P, P, P), mode="reflect").contiguous().to(dev)
pos = (torch.nonzero(torch.nn.functional.pad(mask[0, 0], (P, P, P, P)) == 1)).int().contiguous().to(dev)
mc = pos.shape[0]
out = torch.empty(mc, 25, 25, device=dev)
vp = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
_lib.call("ssl_b200_compute_similarity", vp(img_pad), vp(pos), vp(out), mc, 25, 9, img_pad.shape[1], img_pad.shape[2], 3, st)
gi = torch.zeros_like(img_pad)
_lib.call("ssl_b200_compute_similarity_backward", vp(img_pad), vp(torch.randn(mc, 625, device=dev)), vp(pos), vp(gi), mc, 25, 9,
          img_pad.shape[1], img_pad.shape[2], 3, st)
# host entry
loss_h, grad_h, n = ssl_b200.ssl_step_host(*synth.make_case(2, 48, 40, seed=1, density=0.1), 11, 5)
torch.cuda.synchronize()
print("ok drop-in + host", mc, n)
