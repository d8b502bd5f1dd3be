import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphconvgeo_b200 import ops
dev="cuda"
M=int(sys.argv[1]); N=int(sys.argv[2]); K=int(sys.argv[3]); mode=sys.argv[4]
A = ops.alloc_mat(M, K, dev); A.normal_()
B = ops.alloc_mat(K, N, dev); B.normal_()
out = ops.alloc_mat(M, N, dev)
ops.gemm(A, B, out=out, mode=mode)
torch.cuda.synchronize()
ref = A[:1000].double() @ B.double()
print("ok", M, N, K, mode, (out[:1000].double()-ref).abs().max().item())
