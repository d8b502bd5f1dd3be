import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphconvgeo_b200 import ops
dev="cuda"
def bench(M, N, K, ta, tb, mode, reps=5):
    A = ops.alloc_mat(*((K, M) if ta else (M, K)), dev); A.normal_()
    B = ops.alloc_mat(*((N, K) if tb else (K, N)), dev); B.normal_()
    out = ops.alloc_mat(M, N, dev)
    for _ in range(2): ops.gemm(A, B, out=out, transA=ta, transB=tb, mode=mode)
    ts=[]
    for _ in range(reps):
        s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        s.record(); ops.gemm(A, B, out=out, transA=ta, transB=tb, mode=mode); e.record(); e.synchronize(); ts.append(s.elapsed_time(e))
    ms=float(np.min(ts))
    print("dbg=%s %-7s M=%-6d N=%-6d K=%-6d  %.3f ms  %.1f TFLOP/s" % (os.environ.get("GCG_TC_DEBUG","0"),mode,M,N,K,ms,2.0*M*N*K/ms/1e9), flush=True)
bench(151552, 1024, 608, 0, 1, "tf32")
bench(151552, 1024, 608, 0, 0, "tf32")
bench(151552, 1024, 32, 0, 1, "tf32")
