import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphconvgeo_b200 import ops
dev="cuda"
M,N,K=151552,1024,32
A = ops.alloc_mat(M, K, dev); A.normal_()
B = ops.alloc_mat(N, K, dev); B.normal_()
out = ops.alloc_mat(M, N, dev)
for _ in range(3): ops.gemm(A, B, out=out, transB=True, mode="tf32")
torch.cuda.synchronize(); print("done")
