import os,sys
sys.path.insert(0,".")
import torch, ssl_b200
from ssl_b200 import synth
dev=torch.device("cuda:0")
B=int(os.environ.get("B",16))
sr,gt,mask=synth.make_case(B,256,256,seed=1,density=float(os.environ.get("RHO",0.114)))
sr,gt,mask=sr.to(dev),gt.to(dev),mask.to(dev)
x=sr.clone().requires_grad_(True)
for _ in range(2):
    x.grad=None
    ssl_b200.ssl(x,gt,mask,25,9,0.004,True,path=os.environ.get("SSL_PATH","plane")).backward()
torch.cuda.synchronize()
