#!/usr/bin/env python
"""A_hat.H at the bench's own operands (Twitter-World / -US shaped, nodes reordered by region label exactly as
MLPCONV.prepare does): the register-gather kernel against every streaming variant (gcg_spmm_stream.cu), span
sizes and block x panel schedules.  One JSON line per cell; CUDA events, same process, same box.

    python scripts/spmm_stream_sweep.py --workload twitter-world --F 600 > gpurun_out/stream_sweep.jsonl
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphconvgeo_b200 import _lib, ops, synth  # noqa: E402
from graphconvgeo_b200.sparse import l2_schedule  # noqa: E402


def timeit(fn, reps=5, warm=2):
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
    return float(np.mean(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="twitter-world")
    ap.add_argument("--F", type=int, nargs="+", default=[600])
    ap.add_argument("--variants", type=int, nargs="+", default=[1, 2, 3, 4, 5, 6, 7])
    ap.add_argument("--spans", type=int, nargs="+", default=[384])
    ap.add_argument("--near", type=int, nargs="+", default=[0], help="near windows (rows) to try")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--rows-only", action="store_true", help="whole-row schedule only")
    args = ap.parse_args()
    dev = torch.device("cuda")
    L = _lib.lib()
    peak = 6550.4
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    t0 = time.time()
    wl = synth.make_workload_device(args.workload, device=dev, seed=77)
    A0 = wl.A_hat
    n = A0.shape[0]
    order = np.argsort(wl.Y[:n], kind="stable").astype(np.int32)
    inv = np.empty(n, np.int32)
    inv[order] = np.arange(n, dtype=np.int32)
    A = A0.permute(order, col_map=inv)
    del A0, wl
    torch.cuda.empty_cache()
    print(json.dumps({"built_s": round(time.time() - t0, 1), "n": n, "nnz": A.nnz, "plan": A.plan_info()}), flush=True)

    def emit(tag, F, mean, mn, **kw):
        alg = 8 * A.nnz + 4 * (n + 1) + 8 * n * F
        print(json.dumps(dict(tag=tag, F=F, ms=round(mean, 4), ms_min=round(mn, 4), alg_GBps=round(alg / mean / 1e6, 1),
                              frac=round(alg / mean / 1e6 / peak, 4),
                              gather_TBps=round((8 * A.nnz + 4 * A.nnz * F + 4 * n * F) / mean / 1e9, 2), **kw)), flush=True)

    for F in args.F:
        g = torch.Generator(device=dev).manual_seed(1)
        H = ops.alloc_mat(n, F, dev)
        H.copy_(torch.randn(n, F, device=dev, generator=g))
        out = ops.alloc_mat(n, F, dev)
        ref = ops.spmm(A, H, panel_cols=0).clone()
        emit("register-gather (round 1 default)", F, *timeit(lambda: ops.spmm(A, H, out=out, panel_cols=0)))
        f4 = (F + 3) // 4
        vpl = -(-f4 // 32)
        n_full = max(1, -(-f4 // 32))
        scheds = [("whole rows", None)]
        if vpl > 1 and not args.rows_only:
            scheds.append(("interleaved %d x 128-float panels" % n_full, ([0, n], [-n_full])))
            if vpl >= 4:
                scheds.append(("interleaved 2 panels", ([0, n], [-2])))
        rep = []
        for budget in (() if args.rows_only else (32, 48, 80) if not args.quick else (48,)):
            br, bp = l2_schedule(A, F, budget_bytes=budget << 20, report=rep)
            if bp.max() > 1:
                scheds.append(("l2 tiling budget %d MB: %d blocks, panels hist %s" % (budget, len(bp), np.bincount(bp).tolist()), (br, bp)))
        print(json.dumps({"l2_schedule_candidates": rep}), flush=True)
        for sname, sch in scheds:
            A.set_schedule(*(sch if sch is not None else (None, None)))
            for span in args.spans:
                for var, near in [(v_, n_) for v_ in args.variants for n_ in args.near]:
                    L.gcg_spmm_stream_tuning(var, span, near)
                    try:
                        got = ops.spmm(A, H, out=out, panel_cols=-2)
                        ok = bool(torch.equal(got, ref))
                        emit("stream", F, *timeit(lambda: ops.spmm(A, H, out=out, panel_cols=-2)), variant=var, span=span,
                             near=near, schedule=sname, bit_identical=ok)
                    except Exception as e:      # noqa: BLE001
                        print(json.dumps({"tag": "stream", "F": F, "variant": var, "span": span, "near": near, "schedule": sname,
                                          "error": str(e)[:200]}), flush=True)
                    finally:
                        L.gcg_spmm_stream_tuning(0, 0, -1)
        A.set_schedule(None)
        del H, out, ref
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
