#!/usr/bin/env python
"""One configuration of A_hat.H at the bench's operands, a few launches -- the target of ncu captures.

    ncu --set full --clock-control none -k regex:spmm_ -s 4 -c 1 -o gpurun_out/x python scripts/spmm_stream_one.py --variant 2
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphconvgeo_b200 import _lib, ops, synth  # noqa: E402
from graphconvgeo_b200.sparse import l2_schedule  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="twitter-world")
ap.add_argument("--F", type=int, default=600)
ap.add_argument("--variant", type=int, default=0, help="-1: register-gather kernel of round 1")
ap.add_argument("--span", type=int, default=384)
ap.add_argument("--schedule", default="rows", help="rows | inter<k> | l2:<MB>")
ap.add_argument("--reps", type=int, default=6)
ap.add_argument("--near", type=int, default=-1, help="near window in rows (evict_last inside, evict_first outside)")
args = ap.parse_args()
dev = torch.device("cuda")
wl = synth.make_workload_device(args.workload, device=dev, seed=77)
n = wl.A_hat.shape[0]
order = np.argsort(wl.Y[:n], kind="stable").astype(np.int32)
inv = np.empty(n, np.int32)
inv[order] = np.arange(n, dtype=np.int32)
A = wl.A_hat.permute(order, col_map=inv)
del wl
H = ops.alloc_mat(n, args.F, dev)
H.copy_(torch.randn(n, args.F, device=dev, generator=torch.Generator(device=dev).manual_seed(1)))
out = ops.alloc_mat(n, args.F, dev)
if args.schedule.startswith("inter"):
    A.set_schedule([0, n], [-int(args.schedule[5:])])
elif args.schedule.startswith("l2:"):
    A.set_schedule(*l2_schedule(A, args.F, budget_bytes=int(args.schedule[3:]) << 20))
panel = 0 if args.variant < 0 else -2
_lib.lib().gcg_spmm_stream_tuning(max(args.variant, 0), args.span, args.near)
ts = []
for _ in range(args.reps):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    ops.spmm(A, H, out=out, panel_cols=panel)
    e.record()
    e.synchronize()
    ts.append(s.elapsed_time(e))
print("variant %d span %d schedule %s: ms %s" % (args.variant, args.span, args.schedule, " ".join("%.3f" % t for t in ts)))
