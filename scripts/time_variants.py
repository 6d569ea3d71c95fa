"""Dev helper: per-stage milliseconds of the benchmark step for each library build in build/exp/ (kernel
experiments compiled with -DSSLB_EXPERIMENT_*; results of those builds are NOT correct, only their timing)."""
import glob, os, subprocess, sys
CODE = r'''
import torch, ssl_b200
from ssl_b200 import synth, _lib
sr,gt,mask=synth.make_case(16,256,256,seed=1,density=0.114)
x=sr.cuda().requires_grad_(True); g=gt.cuda(); m=mask.cuda(); n=int(mask.sum())
for i in range(3):
    x.grad=None; l=ssl_b200.ssl(x,g,m,25,9,max_edges=n,parity="global"); l.backward()
_lib.profile_enable(True)
for i in range(5):
    x.grad=None; l=ssl_b200.ssl(x,g,m,25,9,max_edges=n,parity="global"); l.backward()
torch.cuda.synchronize()
print({k:round(v[0]/5,3) for k,v in _lib.profile_read().items() if k in ("ssg_plane_fwd","ssg_plane_bwd","row_loss","plane_eout","plane_finish")})
'''
libs = [None] + sorted(glob.glob("build/exp/*.so"))
for lib in libs:
    env = dict(os.environ, PYTHONPATH=".")
    if lib: env["SSL_B200_LIB"] = os.path.abspath(lib)
    r = subprocess.run([sys.executable, "-c", CODE], capture_output=True, text=True, env=env)
    print((lib or "default").ljust(40), r.stdout.strip().splitlines()[-1] if r.returncode == 0 else r.stderr[-300:], flush=True)
