#!/usr/bin/env python
"""A_hat.Z at the widths of the multi-GPU column slices (F/P): lane-group shapes of the register-gather kernel
and the streaming variants, same Twitter-World-shaped graph as the bench.  One JSON line per cell."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphconvgeo_b200 import _lib, ops, synth  # noqa: E402


def timeit(fn, reps=7, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.mean(ts))


dev = torch.device("cuda")
L = _lib.lib()
wl = synth.make_workload_device("twitter-world", device=dev, seed=77)
n = wl.A_hat.shape[0]
order = np.argsort(wl.Y[:n], kind="stable").astype(np.int32)
inv = np.empty(n, np.int32)
inv[order] = np.arange(n, dtype=np.int32)
A = wl.A_hat.permute(order, col_map=inv)
del wl
for F in (76, 152, 300):
    H = ops.alloc_mat(n, F, dev)
    H.copy_(torch.randn(n, F, device=dev, generator=torch.Generator(device=dev).manual_seed(1)))
    out = ops.alloc_mat(n, F, dev)
    ref = ops.spmm(A, H, panel_cols=0).clone()
    f4 = (F + 3) // 4
    groups = [(0, 0)] + [(g, v) for g, v in ((4, 5), (8, 3), (16, 2), (8, 2), (4, 3)) if g * v >= f4 and g * v < f4 + 8]
    for g, v in groups:
        L.gcg_spmm_set_group(g, v)
        ok = bool(torch.equal(ops.spmm(A, H, out=out, panel_cols=0), ref))
        ms = timeit(lambda: ops.spmm(A, H, out=out, panel_cols=0))
        L.gcg_spmm_set_group(0, 0)
        print(json.dumps({"F": F, "kernel": "register-gather", "lanes": g, "vpl": v, "ms": round(ms, 4), "bit_identical": ok}), flush=True)
    for var in (1, 2, 3, 5):
        L.gcg_spmm_stream_tuning(var, 0, -1)
        try:
            ok = bool(torch.equal(ops.spmm(A, H, out=out, panel_cols=-2), ref))
            ms = timeit(lambda: ops.spmm(A, H, out=out, panel_cols=-2))
            print(json.dumps({"F": F, "kernel": "stream", "variant": var, "ms": round(ms, 4), "bit_identical": ok}), flush=True)
        finally:
            L.gcg_spmm_stream_tuning(0, 0, -1)
    del H, out, ref
