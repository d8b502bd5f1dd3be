#!/usr/bin/env python
"""tcgen05 GEMM bring-up: correctness against float64 for all operand majors, then timing vs FFMA."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphconvgeo_b200 import _lib, ops

L = _lib.lib()
print("tc available:", L.gcg_gemm_tc_available(), flush=True)
dev = "cuda"


def mk(shape, scale):
    return (torch.randn(*shape, device=dev, dtype=torch.float64) * scale)


def check(M, N, K, ta, tb, mode, **kw):
    A64 = mk((K, M) if ta else (M, K), 1.0 / np.sqrt(K))
    B64 = mk((N, K) if tb else (K, N), 1.0)
    A = ops.alloc_mat(*A64.shape, dev); A.copy_(A64)
    B = ops.alloc_mat(*B64.shape, dev); B.copy_(B64)
    ref = (A.double().T if ta else A.double()) @ (B.double().T if tb else B.double())
    out = ops.gemm(A, B, transA=ta, transB=tb, mode=mode, **kw)
    torch.cuda.synchronize()
    err = (out.double() - ref).abs()
    tol = 4e-5 * ref.abs().max() + 1e-4 * ref.abs()
    rel = (err / (ref.abs() + 1e-3)).max().item()
    ok = bool((err <= tol).all())
    print("%-7s M=%-7d N=%-5d K=%-7d tA=%d tB=%d  max_abs_err=%.3e max_rel=%.3e %s" % (mode, M, N, K, ta, tb, err.max().item(), rel, "OK" if ok else ("FAIL" if mode == "tf32x3" else "(plain tf32: informational)")), flush=True)
    return ok


allok = True
for mode in ("tf32x3", "tf32"):
    for (ta, tb) in ((False, False), (False, True), (True, False), (True, True)):
        for (M, N, K) in ((128, 128, 32), (128, 128, 256), (256, 384, 600), (1000, 930, 300), (77, 130, 45), (4096, 600, 600)):
            ok = check(M, N, K, ta, tb, mode)
            if mode == "tf32x3":
                allok &= ok
# split-K (weight gradient shape)
allok &= check(600, 256, 200000, True, False, "tf32x3")
allok &= check(600, 1024, 100000, True, False, "tf32x3", split_k=5)
print("ALL_OK" if allok else "SOME_FAILED", flush=True)


def bench(M, N, K, ta, tb, mode, reps=5):
    A = ops.alloc_mat(*((K, M) if ta else (M, K)), dev); A.normal_()
    B = ops.alloc_mat(*((N, K) if tb else (K, N)), dev); B.normal_()
    out = ops.alloc_mat(M, N, dev)
    for _ in range(2):
        ops.gemm(A, B, out=out, transA=ta, transB=tb, mode=mode)
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.gemm(A, B, out=out, transA=ta, transB=tb, mode=mode); e.record(); e.synchronize()
        ts.append(s.elapsed_time(e))
    ms = float(np.mean(ts))
    print("bench %-7s M=%-8d N=%-5d K=%-8d tA=%d tB=%d  %.3f ms  %.1f TFLOP/s" % (mode, M, N, K, ta, tb, ms, 2.0 * M * N * K / ms / 1e9), flush=True)


for mode in ("fma", "tf32x3", "tf32"):
    bench(450000, 600, 600, False, False, mode)      # H.W
    bench(450000, 600, 256, False, True, mode)       # dZ.W^T
    bench(600, 256, 450000, True, False, mode)       # H^T.dZ
    bench(1400000, 1024, 600, False, False, mode)
