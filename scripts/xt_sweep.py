#!/usr/bin/env python
"""X^T.dZ1 at Twitter-World shape in isolation: document-block size, heavy-row threshold, kernel
family of the blocks, with / without the dense head.  Each configuration runs in a fresh process (the knobs are
read from the environment); this script is the worker:  python scripts/xt_sweep.py  -> one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphconvgeo_b200 import ops, synth  # noqa: E402
from graphconvgeo_b200.lasagne_layers import _x_product, _xt_product  # noqa: E402


class _L:
    pass


dev = torch.device("cuda")
wl = synth.make_workload_device(os.environ.get("WORKLOAD", "twitter-world"), device=dev, seed=77)
X = wl.X
n = X.shape[0]
order = np.argsort(wl.Y[:n], kind="stable").astype(np.int32)
X = X.permute(order)
X.long_row_threshold = 1024
F = wl.hidden
del wl
g = torch.Generator(device=dev).manual_seed(1)
dZ = ops.alloc_mat(n, F, dev)
dZ.copy_(torch.randn(n, F, device=dev, generator=g) * 1e-3)
W = torch.randn(X.shape[1], F, device=dev, generator=g) * 0.01
dW = torch.empty_like(W)
z = ops.alloc_mat(n, F, dev)
layer = _L()


def timed(fn, reps=4):
    fn()
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


t0 = time.time()
t_xt = timed(lambda: _xt_product(layer, X, dZ, dW))
t_xw = timed(lambda: _x_product(layer, X, W, z))
br = layer._xt_blocked[1]
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("GCG_")}, "xt_ms": round(t_xt, 3),
                  "xw_ms": round(t_xw, 3), "n_blocks": len(br.blocks), "n_heavy": br.n_heavy,
                  "heavy_nnz_fraction": round(br.heavy_nnz_fraction, 4), "light_nnz": br.light.nnz,
                  "head": None if getattr(layer, "_x_head", None) is None else layer._x_head[1].k_head,
                  "prep_s": round(time.time() - t0, 1)}), flush=True)
