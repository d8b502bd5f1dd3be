import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphconvgeo_b200 import ops
dev="cuda"
def bench(M, N, K, ta, tb, mode, reps=5, split=1):
    A = ops.alloc_mat(*((K, M) if ta else (M, K)), dev); A.normal_()
    B = ops.alloc_mat(*((N, K) if tb else (K, N)), dev); B.normal_()
    out = ops.alloc_mat(M, N, dev)
    for _ in range(2): ops.gemm(A, B, out=out, transA=ta, transB=tb, mode=mode, split_k=split)
    ts=[]
    for _ in range(reps):
        s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        s.record(); ops.gemm(A, B, out=out, transA=ta, transB=tb, mode=mode, split_k=split); e.record(); e.synchronize(); ts.append(s.elapsed_time(e))
    ms=float(np.min(ts))
    print("%-7s M=%-6d N=%-6d K=%-6d tA=%d tB=%d  %.3f ms  %.1f TFLOP/s" % (mode,M,N,K,ta,tb,ms,2.0*M*N*K/ms/1e9), flush=True)
for mode in ("tf32","tf32x3"):
    for ta,tb in ((0,1),(0,0),(1,0),(1,1)):
        bench(4736, 4096, 8192, ta, tb, mode)      # 37*32 = 1184 tiles = 8 per SM
bench(18944, 1024, 600, 0, 0, "tf32")
bench(18944, 1024, 608, 0, 1, "tf32")
bench(151552, 1024, 608, 0, 1, "tf32")
bench(151552, 1024, 2048, 0, 1, "tf32")
