#!/usr/bin/env python
"""LRU model of the L2 traffic of A_hat.H on the Twitter-World-shaped graph (host only, no GPU).

Generates the benched graph's STRUCTURE with the bench's own generators (synth.city_locations, synth._torch_graph on
the CPU generator: same distribution, not the same random stream as the CUDA generator the bench uses), orders the
nodes by kd-tree region exactly as MLPCONV.prepare does, lays the gathers out in the order the streaming kernel
issues them (spans of 256 consecutive non-zeros, 148 SMs x 16 warps = 2,368 spans in flight, round-robin), and
counts the misses of an LRU cache of whole H rows (2,400 B at F = 600) for several capacities.

  python scripts/l2_model.py [--workload twitter-world] [--out profiles/r02_l2_model.json]
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def lru_lib():
    so = os.path.join(ROOT, "scripts", "_l2_lru.so")        # git-ignored (*.so)
    src = os.path.join(ROOT, "scripts", "l2_lru.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src])
    lib = ctypes.CDLL(so)
    lib.lru_misses.restype = ctypes.c_int64
    lib.lru_misses.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32]
    return lib


def interleave(stream, span=256, in_flight=148 * 16):
    """gathers of `in_flight` consecutive spans issued round-robin (every warp advances at the same rate)"""
    g = span * in_flight
    n_full = (len(stream) // g) * g
    head = stream[:n_full].reshape(-1, in_flight, span).transpose(0, 2, 1).reshape(-1)
    return np.ascontiguousarray(np.concatenate([head, stream[n_full:]]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="twitter-world")
    ap.add_argument("--F", type=int, default=600)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    from graphconvgeo_b200 import synth
    cfg = dict(synth.WORKLOADS[args.workload])
    n_train = cfg["n_train"]
    n = n_train + cfg["n_dev"] + cfg["n_test"]
    t0 = time.time()
    locs, city = synth.city_locations(n, cfg["n_cities"], 77)
    gen = torch.Generator(device="cpu").manual_seed(77)
    ip, ix = synth._torch_graph(n, cfg["avg_deg"], gen, torch.device("cpu"), city=torch.from_numpy(city))
    ip, ix = ip.numpy().astype(np.int64), ix.numpy().astype(np.int64)
    y_train, y_other, _ = synth.assign_classes(locs[:n_train], locs[n_train:], cfg["bucket"])
    Y = np.concatenate([y_train, y_other])
    print("graph + labels in %.0f s: n=%d nnz(adj)=%d regions=%d" % (time.time() - t0, n, len(ix), Y.max() + 1), flush=True)

    def stream_for(order):
        """column ids of A_hat (adjacency + self loops) in row-major order of the permuted matrix"""
        inv = np.empty(n, np.int64)
        inv[order] = np.arange(n)
        rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(ip))
        r2 = np.concatenate([inv[rows], np.arange(n, dtype=np.int64)])        # + self loops
        c2 = np.concatenate([inv[ix], np.arange(n, dtype=np.int64)])
        k = np.argsort(r2 * n + c2, kind="stable")
        return c2[k].astype(np.int32)

    row_bytes = 4 * args.F
    lib = lru_lib()
    sizes = np.bincount(city, minlength=cfg["n_cities"])
    res = {"workload": args.workload, "n": int(n), "nnz_A": int(len(ix) + n), "F": args.F, "row_bytes": row_bytes,
           "largest_cities_nodes": sorted((int(s) for s in sizes), reverse=True)[:8],
           "note": "CPU-generator instance of the bench's graph distribution; LRU over whole H rows", "orders": {}}
    compulsory = n * row_bytes + 8 * (len(ix) + n) + 4 * (n + 1)
    for name, order in (("by region label (MLPCONV.prepare)", np.argsort(Y, kind="stable")),
                        ("by city (the generator's communities)", np.argsort(city, kind="stable")),
                        ("original node ids", np.arange(n))):
        st = stream_for(order)
        entry = {}
        for sched, s in (("2,368 spans in flight, round-robin", interleave(st)), ("one row after the other", st)):
            rows = {}
            for mb in (24, 48, 63, 96, 126):
                cap = int(mb * 2 ** 20 // row_bytes)
                t1 = time.time()
                miss = int(lib.lru_misses(s.ctypes.data, len(s), n, cap))
                gather_gb = miss * row_bytes / 1e9
                total_gb = gather_gb + (8 * len(st) + 4 * (n + 1)) / 1e9 + n * row_bytes / 1e9   # + CSR + output write
                rows["%d MB" % mb] = {"misses": miss, "hit_rate": 1 - miss / len(s), "gather_GB": round(gather_gb, 2),
                                      "dram_GB_with_csr_and_output": round(total_gb, 2)}
                print("%-40s %-36s L2 %3d MB: hit %.3f  gathers %.1f GB  total %.1f GB  (%.0f s)"
                      % (name, sched, mb, 1 - miss / len(s), gather_gb, total_gb, time.time() - t1), flush=True)
            entry[sched] = rows
        res["orders"][name] = entry
    res["compulsory_GB"] = round(compulsory / 1e9, 2)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)
        print("wrote", args.out)


if __name__ == "__main__":
    main()
