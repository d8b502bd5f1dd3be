#!/usr/bin/env python
"""Isolated A_hat.H SpMM sweep (BASELINE config 5): power-law graphs x feature widths x kernel
variants, timed with CUDA events.  One JSON line per cell on stdout.

    python scripts/spmm_sweep.py --n 1000000 --deg 5 20 --F 64 256 600 --graph chunglu community
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
from graphconvgeo_b200.sparse import CSRMatrix, _np_ptr  # noqa: E402


def build_graph(n, deg, kind, dev, seed=77, comm_size=2400):
    """kinds: chunglu (no locality at all) | community (region-sized communities of ~comm_size nodes, 80 % of the
    edges inside; ids drawn at random, so the node order carries no locality until it is reordered) |
    community-sorted (the same communities with contiguous node ids = the graph after reordering by region)."""
    from graphconvgeo_b200.sparse import build_ahat_device
    gen = torch.Generator(device=dev).manual_seed(seed)
    city = None
    if kind.startswith("community"):
        k = max(1, n // comm_size)
        if kind == "community-sorted":
            city = torch.arange(n, device=dev) // comm_size
        else:
            city = torch.randint(0, k, (n,), generator=gen, device=dev)
    ip, ix = synth._torch_graph(n, deg, gen, dev, city=city)
    A = build_ahat_device(ip, ix, n)          # the product's device builder (bit-identical to the host one)
    return A, (city.cpu().numpy() if city is not None else None)


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
    ap.add_argument("--n", type=int, nargs="+", default=[1000000])
    ap.add_argument("--deg", type=int, nargs="+", default=[20])
    ap.add_argument("--F", type=int, nargs="+", default=[600])
    ap.add_argument("--graph", nargs="+", default=["chunglu"])
    ap.add_argument("--panel", type=int, nargs="+", default=[0])
    ap.add_argument("--reorder", nargs="+", default=["none"])
    ap.add_argument("--thr", type=int, default=256)
    ap.add_argument("--tune", nargs="+", default=["0,0"], help="u,minb pairs for gcg_spmm_set_tuning")
    args = ap.parse_args()
    dev = torch.device("cuda")
    peak = 6550.4
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    for n in args.n:
        for deg in args.deg:
            for kind in args.graph:
                t0 = time.time()
                A0, comm = build_graph(n, deg, kind, dev)
                for ro in args.reorder:
                    A = A0
                    if ro != "none":
                        if ro == "community" and comm is not None:
                            order = np.argsort(comm, kind="stable").astype(np.int32)
                        elif ro == "degree":
                            order = np.argsort(-np.diff(A0._host_arrays()[0]), kind="stable").astype(np.int32)
                        else:
                            continue
                        inv = np.empty(n, np.int32)
                        inv[order] = np.arange(n, dtype=np.int32)
                        A = A0.permute(order, col_map=inv)
                    A.long_row_threshold = args.thr
                    info = A.plan_info()
                    for F in args.F:
                        if 2 * n * F * 4 > 100e9:
                            continue
                        B = torch.randn(n, F, device=dev)
                        Bm = ops.alloc_mat(n, F, dev)
                        Bm.copy_(B)
                        del B
                        out = ops.alloc_mat(n, F, dev)
                        alg = 8 * A.nnz + 4 * (n + 1) + 8 * n * F
                        for panel in args.panel:
                          for tune in args.tune:
                            tu, tm = (int(x) for x in tune.split(","))
                            _lib.lib().gcg_spmm_set_tuning(tu, tm)
                            pc = None if panel == -9 else panel          # -9: ops.spmm's own choice of kernel
                            mean, mn = timeit(lambda: ops.spmm(A, Bm, out=out, panel_cols=pc))
                            _lib.lib().gcg_spmm_set_tuning(0, 0)
                            print(json.dumps({"tune": tune, "n": n, "deg": deg, "graph": kind, "reorder": ro, "F": F, "panel": panel,
                                              "kernel": {-9: "auto:" + ("stream" if ops.auto_panel_cols(n, F, A.nnz) == -2 else "gather"),
                                                         -2: "stream", -1: "bulk-copy", 0: "gather"}.get(panel, "gather panel %d" % panel),
                                              "nnz": A.nnz, "max_deg": info["max_degree"], "n_long": info["n_long_rows"],
                                              "ms": round(mean, 4), "ms_min": round(mn, 4),
                                              "alg_GBps": round(alg / mean / 1e6, 1), "frac_of_measured_peak": round(alg / mean / 1e6 / peak, 4),
                                              "gather_GBps": round((8 * A.nnz + 4 * A.nnz * F + 4 * n * F) / mean / 1e6, 1)}), flush=True)
                        del Bm, out
                del A0
                torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
