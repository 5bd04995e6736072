import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import torch
from weatherbench2_b200 import _lib
ctx=_lib.Context(0)
NLAT,NLON=721,1440
nfield=390
x=torch.randn((nfield,NLAT,NLON),device='cuda',dtype=torch.float32)
nk=NLON//2+1
scale=np.cos(np.deg2rad(np.linspace(-90,90,NLAT)))*4.0e7
acc=torch.zeros((13,NLAT,nk),device='cuda',dtype=torch.float32)
torch.cuda.synchronize()
for _ in range(3):
  ctx.zonal_spectrum(x.data_ptr(),nfield,NLAT,NLON,scale,acc.data_ptr(),True,13)
ctx.synchronize()
