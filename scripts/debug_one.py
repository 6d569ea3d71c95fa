import sys, torch, ssl_b200
from ssl_b200 import synth
b,h,w,ks,kw,rho = eval(sys.argv[1]); grad = sys.argv[2] == "1"
sr, gt, mask = synth.make_case(b,h,w,seed=3,density=rho)
x = sr.cuda().requires_grad_(grad)
loss = ssl_b200.ssl(x, gt.cuda(), mask.cuda(), ks, kw, path="plane")
if grad: loss.backward()
torch.cuda.synchronize()
print("ok", float(loss))
