"""One config-2 plane forward (for ncu)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssl_b200
from ssl_b200 import _lib, synth
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
dev = torch.device("cuda:0")
vp = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())
B = int(os.environ.get("B", 16))
sr, gt, mask = synth.make_case(B, 256, 256, seed=1, density=0.114)
sr, gt, mask = sr.to(dev), gt.to(dev), mask.to(dev)
el = ssl_b200.build_edge_list(mask)
n = el.count()
rows = torch.empty(n, 625, device=dev); rows2 = torch.empty_like(rows)
nb = int(_lib.load().ssl_b200_plane_rows_workspace_bytes(B, 256, 256, 25, 9, n))
ws = torch.empty(nb, dtype=torch.uint8, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(2):
    _lib.call("ssl_b200_plane_rows_forward", vp(sr), vp(gt), 0, B, 3, 256, 256, vp(el.edges), vp(el.counts), n, 25, 9,
              vp(rows), vp(rows2), vp(ws), nb, st)
torch.cuda.synchronize()
print("done", n)
