import sys, numpy as np, traceback
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import scipy.sparse as sp
import torch
import spral_b200 as sb
from spral_b200 import matrices as M, _lib
rng = np.random.default_rng(3)
n=40
A = rng.uniform(-1,1,(n,n)); A=(A+A.T)/2
A[:,5]=0; A[5,:]=0; A[:,17]=0; A[17,:]=0
n_,ptr,row,val = M._lower_csc_keep_zeros(sp.csc_matrix(A))
ak = sb.analyse(n_,ptr,row,order=np.arange(1,n+1,dtype=np.int32))
print("nnodes", ak.analysis.nnodes, "maxfront", ak.analysis.maxfront)
fk = sb.factor(ak, False, val)
print("action=T", fk.inform, fk.numeric[0].stats.as_dict())
torch.cuda.synchronize(); print("sync ok")
opt = _lib.Options.default(); opt.action = False
fk2 = sb.factor(ak, False, val, options=opt)
print("action=F", fk2.inform)
torch.cuda.synchronize(); print("sync ok 2")
x = torch.ones(10).cuda(); print(x.sum().item())
