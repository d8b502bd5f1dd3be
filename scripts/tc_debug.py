import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphconvgeo_b200 import ops
dev="cuda"
torch.set_printoptions(linewidth=200, precision=1, sci_mode=False)
def run(A, B, ta, tb, mode="tf32"):
    Ad = ops.alloc_mat(*A.shape, dev); Ad.copy_(A)
    Bd = ops.alloc_mat(*B.shape, dev); Bd.copy_(B)
    out = ops.gemm(Ad, Bd, transA=ta, transB=tb, mode=mode)
    torch.cuda.synchronize()
    return out
M=N=128; K=32
ones = torch.ones(M, K)
# NN, B[k][n] = n
B = torch.arange(N, dtype=torch.float32)[None,:].repeat(K,1)
o = run(ones, B, False, False)
print("NN B=n   row0[:16]", o[0,:16].cpu().tolist(), " expect 32*n"); print("   row0[32:40]", o[0,32:40].cpu().tolist(), "row5[100:104]", o[5,100:104].cpu().tolist())
B = torch.arange(K, dtype=torch.float32)[:,None].repeat(1,N)
o = run(ones, B, False, False)
print("NN B=k   row0[:8]", o[0,:8].cpu().tolist(), " expect 496")
# one-hot B: B[k0][n0]=1
for (k0,n0) in ((0,0),(1,0),(0,1),(8,0),(0,32),(9,33)):
    B = torch.zeros(K,N); B[k0,n0]=1
    A = torch.arange(M*K, dtype=torch.float32).reshape(M,K) % 256
    o = run(A, B, False, False)
    nz = torch.nonzero(o.cpu())
    print("NN onehot B[%d][%d]: nonzero count %d, first %s ; expect column %d = A[:,%d]" % (k0,n0,len(nz), nz[:3].tolist(), n0, k0), "col vals", o[:4,n0].cpu().tolist(), "A", A[:4,k0].tolist())
# K-major reference case (NT)
B = torch.arange(N, dtype=torch.float32)[:,None].repeat(1,K)   # stored [N][K], B^T[k][n]=n
o = run(ones, B, False, True)
print("NT B=n   row0[:8]", o[0,:8].cpu().tolist())
