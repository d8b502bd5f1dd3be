#!/usr/bin/env python
"""bench.py -- GCN fwd+bwd epoch time and A_hat.H SpMM HBM GB/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload twitter-world|twitter-us|geotext]
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on host cores
    torchrun ... bench.py --gpus N ...        # one rank per GPU, row-partitioned A_hat

A "step" is one epoch: one f_train call = full-graph forward + backward + Adam step
(mlpconv.py:293-295) of the 3-layer highway GCN on a seeded synthetic workload of the named
shape.  `value` is device time per epoch with inputs resident in HBM; `e2e` re-uploads the
epoch's host inputs (X CSR, labels, target indices -- what f_train(X, Y, idx) receives) from
pinned memory and reads loss/acc back every step.  `roofline` is the A_hat.H SpMM kernel
(algorithmic bytes 8*nnz + 4*(N+1) + 8*N*F over its CUDA-event time).  One JSON line on stdout.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock / throttle reasons sampled every 200 ms while the timed region runs."""

    def __init__(self, index=0):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:            # pragma: no cover
            self.nv = None
            log("clock sampler unavailable:", e)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------ CPU baseline
def host_workload(workload, log_fn=log):
    """The bench's own workload (same seeds, same N / V / nnz / regions as the GPU arm) as host arrays, built
    WITHOUT libgcg.so: raw adjacency / TF-IDF / coordinates from the seeded torch generators (on the GPU when
    there is one -- data plumbing -- else on the CPU), then A_hat by the oracle's restatement of
    tensormain.py:170-180,221 and the labels by the oracle's kd-tree / nearest-median (kdtree.py, data.py:399-421)."""
    import scipy.sparse as sp
    import torch
    from graphconvgeo_b200 import synth
    from oracle import gcn_oracle as go
    from oracle import kdtree_oracle as ko
    t0 = time.time()
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    cfg, n, locs, (ip, ix), (xip, xix, xv) = synth.make_raw_device(workload, device=dev, seed=77)
    adj = sp.csr_matrix((np.ones(ix.numel(), np.float64), ix.cpu().numpy(), ip.cpu().numpy()), shape=(n, n))
    X = sp.csr_matrix((xv.cpu().numpy(), xix.cpu().numpy(), xip.cpu().numpy()), shape=(n, cfg["vocab"]))
    del ip, ix, xip, xix, xv
    if dev == "cuda":
        torch.cuda.empty_cache()
    t1 = time.time()
    A = go.build_ahat(adj)
    del adj
    t2 = time.time()
    n_train = cfg["n_train"]
    y_train, n_leaves = ko.kdtree_labels(locs[:n_train], cfg["bucket"])
    med = ko.cluster_medians(locs[:n_train], y_train)
    y_other = ko.nearest_median_labels(locs[n_train:], med)
    Y = np.concatenate([y_train, y_other]).astype(np.int64)
    t3 = time.time()
    log_fn("[reference] workload %s on the host: raw %.1fs (torch on %s), A_hat (oracle) %.1fs, labels (oracle) %.1fs"
           % (workload, t1 - t0, dev, t2 - t1, t3 - t2))
    meta = dict(n=n, vocab=cfg["vocab"], hidden=cfg["hidden"], regions=int(Y.max()) + 1, nnz_A=int(A.nnz),
                nnz_X=int(X.nnz), max_degree=int(np.diff(A.indptr).max()))
    return X, A, Y, np.arange(n_train, dtype=np.int32), meta


# timed epochs the reference arm runs per workload: one Twitter-World epoch of the oracle takes ~100 s of host time
REF_EPOCH_CAP = {"twitter-world": (1, 1), "twitter-us": (1, 3), "geotext": (3, 20), "tiny": (3, 20)}    # (warm-up, timed)


def cpu_reference_epochs(workload, n_layers, highway, steps, warmup):
    """The reference's CPU path -- the oracle port (scipy csr@dense on 1 thread, what Theano's S.dot dispatches
    to, + multi-threaded BLAS) -- on the FULL workload: whole epochs of mlpconv.py:293-295, timed one by one."""
    from oracle import gcn_oracle as go
    threads = os.cpu_count()
    X, A, Y, train_idx, meta = host_workload(workload)
    # peak resident set of a lean oracle epoch, measured: 53 GB at Twitter-World = 9.3 [N, max(h, C)] float32
    # arrays; refuse rather than swap / be OOM-killed
    need = 10.0 * meta["n"] * max(meta["hidden"], meta["regions"]) * 4
    try:
        import psutil
        avail = float(psutil.virtual_memory().available)
    except Exception:
        avail = float("inf")
    if need > 0.95 * avail and os.environ.get("GCG_REF_FORCE") != "1":
        print(json.dumps({"impl": "reference", "unavailable": "host memory: the %s oracle epoch needs ~%.0f GB, "
                          "%.0f GB available" % (workload, need / 1e9, avail / 1e9)}), flush=True)
        raise SystemExit(0)
    rng = np.random.RandomState(0)
    params = go.init_params(rng, X.shape[1], meta["hidden"], meta["regions"], n_layers, highway)
    net = go.GCNOracle(X, A, n_layers, highway, (1e-6, 1e-6))
    y = Y[train_idx].astype(np.int32)
    st = go.AdamState(params)
    cap_w, cap_s = REF_EPOCH_CAP[workload]
    warmup, steps = min(warmup, cap_w), max(1, min(steps, cap_s))
    times, all_times, loss = [], [], None
    budget_s = float(os.environ.get("GCG_REF_BUDGET_S", "900"))       # the arm must end within minutes
    t_arm = time.perf_counter()
    i = -1
    while True:
        i += 1
        if i >= warmup + steps:
            break
        if i > 0 and (time.perf_counter() - t_arm) * (i + 1.0) / i > budget_s:
            # the next epoch would overrun: what has run so far is the measurement (the first epoch then counts
            # as timed, not as warm-up: scipy / BLAS have no compile or cache warm-up to exclude)
            if not times:
                times, warmup = list(all_times), 0
            break
        t0 = time.perf_counter()
        loss, acc, grads, _c = net.loss_and_grads(params, train_idx, y, lean=True)
        go.adam_step(params, grads, st)
        dt = (time.perf_counter() - t0) * 1e3
        del grads, _c
        import resource
        log("[reference] epoch %d: %.1f ms  loss %.5f acc %.4f  (peak RSS %.1f GB)"
            % (i, dt, float(loss), acc, resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6))
        all_times.append(dt)
        if i >= warmup:
            times.append(dt)
    desc = ("%d whole epoch(s) (after %d warm-up) of the oracle port (scipy csr@dense on 1 thread, as under Theano, + BLAS on "
            "%d threads) on the full %s workload: %d nodes, vocab %d, %d regions, nnzA %d, nnzX %d"
            % (len(times), warmup, threads, workload, meta["n"], meta["vocab"], meta["regions"], meta["nnz_A"], meta["nnz_X"]))
    return float(np.mean(times)), desc, threads, len(times), warmup, meta


def cpu_baseline_sample(m, wl, n_layers, highway, row_fraction=1.0 / 32, seed=5):
    """cpu_baseline of the GPU arm: a BOUNDED sample of the same full-size workload.  The epoch's row-parallel
    products (X.W1, A_hat.H, the dense projections and their gradients) are timed with the oracle's own routines
    (scipy csr@dense, BLAS) on a random ``row_fraction`` of the output rows against FULL-size operands (all N
    gathered rows, the whole vocabulary, all regions) and scaled by 1/row_fraction; Adam is timed on the full
    parameter set.  Same N, V, C and non-zeros as the GPU arm; `--impl reference` runs whole epochs."""
    import scipy.sparse as sp
    from oracle import gcn_oracle as go
    t_start = time.perf_counter()
    X = m.Xd.to_scipy()                       # the GPU arm's own (region-reordered) inputs, copied to the host
    A = m.l_hid1.H.to_scipy()
    N, V = X.shape
    h, C = wl.hidden, wl.n_classes
    rng = np.random.RandomState(seed)
    rows = np.sort(rng.choice(N, size=max(1, int(N * row_fraction)), replace=False))
    Xr, Ar = X[rows], A[rows]
    XrT = sp.csr_matrix(Xr.T)
    params = go.init_params(np.random.RandomState(0), V, h, C, n_layers, highway)
    W1, Wout = params[0], params[-2]
    Hfull = rng.standard_normal((N, h)).astype(np.float32)
    Hr = np.ascontiguousarray(Hfull[rows])
    Gr = rng.standard_normal((len(rows), C)).astype(np.float32)
    t = {}

    def timed(key, fn, reps=1):
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        t[key] = (time.perf_counter() - t0) / reps
        return r
    timed("X.W1", lambda: Xr @ W1)
    timed("A_hat.H (F=h)", lambda: Ar @ Hfull)
    timed("X^T.dZ1", lambda: XrT @ Hr)
    timed("H.W (h x h)", lambda: np.dot(Hr, params[2] if n_layers > 2 else W1[:h, :h].copy()))
    timed("H^T.dZ (h x h)", lambda: np.dot(Hr.T, Hr))
    timed("H.W_out", lambda: np.dot(Hr, Wout))
    timed("H^T.dP_out", lambda: np.dot(Hr.T, Gr))
    timed("dP_out.W_out^T", lambda: np.dot(Gr, Wout.T))
    grads = [np.zeros_like(p) for p in params]
    st = go.AdamState(params)
    timed("adam (all parameters)", lambda: go.adam_step(params, grads, st))
    n_hid = n_layers - 2
    gemm_hh = (2 if highway else 1) * n_hid       # H.W (+ H.Wg) forward; twice that many of each kind backward
    scaled = (t["X.W1"] + t["X^T.dZ1"]
              + (2 + 2 * n_hid) * t["A_hat.H (F=h)"]                      # fwd + bwd per hidden conv layer, layer 1 included
              + 2 * t["A_hat.H (F=h)"] * (n_train_rows(m) / float(N))     # output layer propagates the target rows only
              + gemm_hh * (t["H.W (h x h)"] + t["H^T.dZ (h x h)"] + t["H.W (h x h)"])
              + t["H.W_out"] + t["H^T.dP_out"] + t["dP_out.W_out^T"]) / row_fraction + t["adam (all parameters)"]
    desc = ("oracle routines (scipy csr@dense 1 thread + BLAS on %d threads) on a random %.4g of the output rows of the "
            "full %s workload (%d nodes, vocab %d, %d regions, nnzA %d, nnzX %d; operands full size), scaled by %g; "
            "Adam on all parameters; %.1f s of host time" % (os.cpu_count(), row_fraction, wl.name, N, V, C, A.nnz, X.nnz,
                                                             1.0 / row_fraction, time.perf_counter() - t_start))
    return scaled * 1e3, desc, {k: round(v * 1e3, 2) for k, v in t.items()}


def n_train_rows(m):
    return m.ti_train.n


def cpu_mlp_epoch(Xc_host, y_train, hidden, n_classes, batch, n_batches, sample_batches=4):
    """cpu_baseline of the minibatch MLP (used by scripts/smooth_bench.py): the NumPy
    port of mlp.py:267-271 timed on ``sample_batches`` minibatches, scaled to one epoch.  Returns ms."""
    from oracle import mlp_oracle as mo
    params = mo.init_params(np.random.RandomState(0), Xc_host.shape[1], hidden, n_classes)
    net = mo.MLPOracle((1e-6, 1e-6))
    rng = np.random.RandomState(0)
    st = mo.AdamState(params)
    sample_batches = max(1, min(n_batches, sample_batches))
    t0 = time.perf_counter()
    for k, idx in enumerate(mo.iterate_minibatches(Xc_host.shape[0], batch, rng)):
        if k == sample_batches:
            break
        _, _, grads = net.loss_and_grads(params, Xc_host[idx], y_train[idx])
        mo.adam_step(params, grads, st, lr=2e-3)
    per_batch = (time.perf_counter() - t0) / sample_batches
    return per_batch * n_batches * 1e3, sample_batches


def cpu_projection(B, n_targets, budget_pairs=4e6):
    """cpu_baseline of the mention-graph projection (used by scripts/projection_bench.py):
    the plain-Python port of data.py:226-250 on the first nodes whose clique pairs fit ``budget_pairs``,
    scaled to the whole graph by pair count.  Returns (ms for the whole graph, nodes sampled)."""
    from oracle import graph_oracle as gro
    import scipy.sparse as sp
    B = sp.csr_matrix(B)
    tdeg = np.diff(sp.csr_matrix(B[:, :n_targets]).indptr).astype(np.float64)
    pairs = tdeg * tdeg
    cum = np.cumsum(pairs)
    k = int(min(len(cum), max(1, np.searchsorted(cum, budget_pairs) + 1)))
    adj = gro.adjacency_sets(B[:k])
    t0 = time.perf_counter()
    gro.project(adj, n_targets)
    dt = time.perf_counter() - t0
    return dt * cum[-1] / max(cum[k - 1], 1.0) * 1e3, k


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; Theano is not
    installable here) on the box's host cores, on the SAME workload as the GPU arm (same seeds, N, V, regions,
    non-zeros), whole epochs; `steps` / `warmup` report what was actually timed (a Twitter-World epoch takes
    ~100 s of host time, so at most 1 + 1 are run).  Nothing of libgcg.so is loaded on this path."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.random_graph:
        raise SystemExit("--impl reference runs the community graph (the benched configuration) only")
    ms, desc, threads, n_timed, n_warm, meta = cpu_reference_epochs(args.workload, args.layers, bool(args.highway),
                                                                     args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "gcn_fwd_bwd_epoch_ms", "value": ms, "unit": "ms", "n_gpus": args.gpus,
        "steps": n_timed, "warmup": n_warm, "ms_per_step": ms, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(bench_config(args), nodes=meta["n"], vocab=meta["vocab"], hidden=meta["hidden"],
                       regions=meta["regions"], nnz_A=meta["nnz_A"], nnz_X=meta["nnz_X"], max_degree=meta["max_degree"],
                       graph="community"),
        "engine": {"parallelism": "host: scipy csr@dense on 1 thread (as under Theano) + BLAS on %d threads" % threads},
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def bench_config(args):
    return {"workload": "%s-shaped synthetic, %d-layer %sGCN, full-batch fwd+bwd+Adam epoch"
                        % (args.workload, args.layers, "highway " if args.highway else ""),
            "n_layers": args.layers, "highway": bool(args.highway),
            "l2": "flush" if args.workload in ("geotext", "tiny") else "inputs larger than L2"}


# ------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from graphconvgeo_b200 import ops, synth
    from graphconvgeo_b200.mlpconv import MLPCONV

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: libraries (NCCL prints its version banner there) are
    # redirected to stderr for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=150))

    t0 = time.time()
    wl = synth.make_workload_device(args.workload, device=dev, seed=77, community=not args.random_graph)
    log("[rank %d] workload %s built in %.1fs: %s" % (rank, args.workload, time.time() - t0, wl.meta))

    model_kwargs = dict(n_epochs=args.steps, regul_coefs=[1e-6, 1e-6], hidden_layer_size=wl.hidden, drop_out=False,
                        n_layers=args.layers, highway=bool(args.highway), seed=1, device=dev, cuda_graph=not args.no_graph)
    if world > 1:
        from graphconvgeo_b200.dist import DistMLPCONV
        m = DistMLPCONV(partition=args.partition, peer_memory=not args.no_peer_memory, **model_kwargs)
        m.keep_host_inputs = not args.no_parity
    else:
        m = MLPCONV(**model_kwargs)
    t0 = time.time()
    m.prepare(wl.X, wl.train_indices, wl.dev_indices, wl.test_indices, wl.Y, wl.A_hat)
    log("[rank %d] model prepared in %.1fs" % (rank, time.time() - t0))

    flush = None
    if bench_config(args)["l2"] == "flush":
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- warm-up (epoch 1 eager, then graph capture, then replays)
    for _ in range(max(args.warmup, 3)):
        m.f_train()
    barrier()
    ops.launch_count(reset=True)
    launches_per_step = m.launches_per_step if hasattr(m, "launches_per_step") else None

    # ---------------- timed region: exactly K epochs, device time, max over ranks
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    barrier()
    with sampler:
        for s, e in ev:
            if flush is not None:
                flush.fill_(1)
            s.record()
            m.f_train()
            e.record()
        barrier()
    step_ms = [s.elapsed_time(e) for s, e in ev]
    total_ms = float(sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    loss, acc = m.train_results()
    log("[rank %d] epoch %.3f ms (min %.3f max %.3f) loss %.5f acc %.4f" % (rank, ms_per_step, min(step_ms), max(step_ms), loss, acc))

    # ---------------- kernel launches per epoch (counted eagerly, once)
    m_graph, m._graph = m._graph, None
    ops.launch_count(reset=True)
    m._train_step_enqueue()
    torch.cuda.synchronize(dev)
    launches = ops.launch_count(reset=True)
    m._graph = m_graph

    breakdown = op_breakdown(m, dev) if args.breakdown else None       # collective-safe: every rank runs it
    if world > 1 and os.environ.get("GCG_DIST_PROFILE") == "1":
        from graphconvgeo_b200.dist import phase_profile
        phase_profile.acc.clear()
        phase_profile.pending = []
        m._train_step_enqueue()
        rep = phase_profile.report()
        if rank == 0:
            log("---- phases of the feature-sliced propagation over one epoch (rank 0) ----")
            for k, v in rep.items():
                log("  %-10s x%-3d %8.3f ms" % (k, v["calls"], v["ms"]))

    # ---------------- e2e: host inputs re-uploaded every epoch, loss/acc read back
    e2e = measure_e2e(m, args, dev, flush)        # every rank re-uploads its own row block
    if world > 1:
        t = torch.tensor([e2e["value"], float(e2e["h2d_bytes_per_step"])], device=dev, dtype=torch.float64)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        e2e["value"] = float(tmax[0].item())                      # max over ranks
        e2e["h2d_bytes_per_step"] = int(t[1].item())              # summed over ranks

    # ---------------- roofline of the headline kernel: A_hat . H  (F = hidden)
    roof = spmm_roofline(m, wl, dev)          # collective in row-partitioned mode: every rank calls it

    # ---------------- parity of one training step at this size, operation by operation on sampled rows
    parity = None if args.no_parity else parity_block(m, args, rank)     # collective: every rank calls it

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        frac = {"twitter-world": 1.0 / 32, "twitter-us": 1.0 / 16}.get(args.workload, 1.0)
        ms, desc, per_op = cpu_baseline_sample(m, wl, args.layers, bool(args.highway), row_fraction=frac)
        cpu = {"value": ms, "unit": "ms", "cores": os.cpu_count(), "kind": "port", "sample": desc,
               "sample_op_ms": per_op}

    line = {
        "metric": "gcn_fwd_bwd_epoch_ms", "value": ms_per_step, "unit": "ms", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # `config` names the WORKLOAD only and is identical, key for key, in the reference arm's line (the driver
        # compares the two); how this arm executes it is in `engine`
        "config": dict(bench_config(args), nodes=wl.meta["n"], vocab=wl.meta["vocab"], hidden=wl.hidden,
                       regions=wl.n_classes, nnz_A=wl.meta["nnz_A"], nnz_X=wl.meta["nnz_X"],
                       max_degree=wl.meta["max_degree"], graph="community" if wl.meta["community"] else "chung-lu"),
        "engine": dict(parallelism=("rows x%d, A_hat.Z %s" % (world, ("feature-sliced, transposes by %s" % (
                           "peer-memory stores over NVLink" if getattr(getattr(m, "part", None), "peer", None) is not None
                           else "NCCL all-to-all")) if getattr(m, "partition", "") == "feature" else "row blocks + NCCL all-gather"))
                       if world > 1 else "single GPU",
                       gemm_mode=os.environ.get("GCG_GEMM_MODE", "auto"),
                       **({"collectives": ("libgcg.so (gcg_comm_*: NCCL bound by the C ABI)" if getattr(m, "collectives", "torch") == "native"
                                           else "torch.distributed (NCCL)")} if world > 1 else {}),
                       epoch_driver=("gcg_epoch_run (the epoch recorded as a C++ call list, include/gcg.h), captured in a CUDA graph"
                                     if getattr(m, "_program", None) is not None else
                                     ("layer code captured in a CUDA graph" if getattr(m, "_graph", None) is not None
                                      else "layer code, eager"))),
        "clocks": sampler.summary(), "gpu_launches": launches * args.steps,
        "launches_per_epoch": launches, "loss": loss, "acc": acc,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if roof is not None:
        line["roofline"] = roof["roofline"]
        line["spmm"] = roof["detail"]
    alg = algorithmic_table(m, wl, args.layers, bool(args.highway))
    peak_gbs, _src = measured_peaks()
    alg["floor_ms_at_measured_hbm_peak"] = alg["total_bytes"] / (peak_gbs * 1e9) * 1e3
    line["algorithmic"] = alg
    log("---- algorithmic work of one epoch (SURVEY 8d): %.2f GB compulsory, %.2f TFLOP; %.2f ms at %.0f GB/s ----"
        % (alg["total_bytes"] / 1e9, alg["total_flops"] / 1e12, alg["floor_ms_at_measured_hbm_peak"], peak_gbs))
    for r in alg["rows"]:
        log("  %-46s x%-2d %9.3f GB %9.3f TFLOP" % (r["op"], r["count"], r["bytes_each"] / 1e9, r["flops_each"] / 1e12))
    if parity is not None:
        line["parity"] = parity
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if breakdown is not None:
        line["breakdown"] = breakdown
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def algorithmic_table(m, wl, n_layers, highway):
    """compulsory traffic (each array touched once) and flops of one epoch, operation by operation.
    B_spmm(nnz, n_out, n_in, F) = 8*nnz + 4*(n_out+1) + 4*n_in*F + 4*n_out*F."""
    N, V, h, C = wl.meta["n"], wl.meta["vocab"], wl.hidden, wl.n_classes
    nnzA, nnzX = wl.meta["nnz_A"], wl.meta["nnz_X"]
    n_idx = len(wl.train_indices)
    nnz_sub = int(round(nnzA * n_idx / float(N)))                  # A_hat[idx, :]
    pf = bool(getattr(m.l_out, "propagate_first", False))
    bs = lambda nnz, n_out, n_in, F: 8 * nnz + 4 * (n_out + 1) + 4 * n_in * F + 4 * n_out * F
    gemm = lambda M, Nn, K: (4 * (M * K + K * Nn + M * Nn), 2 * M * Nn * K)
    rows = [("X.W1 (a1)", 1, bs(nnzX, N, V, h), 2 * nnzX * h),
            ("A_hat.Z, F=h, fwd (a4)", n_layers - 1, bs(nnzA, N, N, h), 2 * nnzA * h),
            ("A_hat.dP, F=h, bwd (a5)", n_layers - 1 + (1 if pf else 0), bs(nnzA, N, N, h), 2 * nnzA * h),
            ("X^T.dZ1 (a6)", 1, bs(nnzX, V, N, h), 2 * nnzX * h)]
    n_hid = n_layers - 2
    per_hid = (2 if highway else 1)
    rows.append(("H.W / H.Wg, h x h, fwd (a3, a9)", n_hid * per_hid) + gemm(N, h, h))
    rows.append(("H^T.dZ / H^T.dG, h x h (a7)", n_hid * per_hid) + gemm(h, h, N))
    rows.append(("dZ.W^T / dG.Wg^T, h x h (a7)", n_hid * per_hid) + gemm(N, h, h))
    if highway and n_hid:
        rows.append(("highway mix epilogue + backward (a9)", n_hid, (16 + 28) * N * h, 12 * N * h))
    if pf:
        rows.append(("A_hat[idx,:].H, F=h (a3, propagate first)", 1, bs(nnz_sub, n_idx, N, h), 2 * nnz_sub * h))
        rows.append(("(.).W_out, Q^T.dP, dP.W_out^T (a3, a7)", 3) + gemm(n_idx, C, h))
    else:
        rows.append(("H.W_out, H^T.dZ, dZ.W_out^T (a3, a7)", 3) + gemm(N, C, h))
        rows.append(("A_hat[idx,:].Z, F=C (a3)", 1, bs(nnz_sub, n_idx, N, C), 2 * nnz_sub * C))
        rows.append(("A_hat.dP, F=C, bwd (a5)", 1, bs(nnzA, N, N, C), 2 * nnzA * C))
    rows.append(("softmax + CE + argmax + dLogits (a10)", 1, 8 * n_idx * C, 6 * n_idx * C))
    n_par = sum(p.numel() for p in m.params)
    rows.append(("Adam + elastic net (a11, a12)", 1, 28 * n_par, 12 * n_par))
    tot_b = sum(c * b for _, c, b, _f in rows)
    tot_f = sum(c * f for _, c, _b, f in rows)
    return {"rows": [{"op": o, "count": c, "bytes_each": int(b), "flops_each": int(f)} for o, c, b, f in rows],
            "total_bytes": int(tot_b), "total_flops": int(tot_f)}


def parity_block(m, args, rank):
    """One eager forward + one eager f_train of the benched model checked against the oracle on sampled rows
    (oracle/sampled_parity.py): max over all operations of |gpu - oracle| / (1e-6 + 1e-4*|oracle|)."""
    from oracle import sampled_parity
    t0 = time.time()
    g, m._graph = getattr(m, "_graph", None), None
    try:
        Xh, Ah = m.host_inputs()
        rep = sampled_parity.check_training_step(m, Xh, Ah, n_rows=args.parity_rows, seed=3,
                                                 log=log if rank == 0 else None)
    finally:
        m._graph = g
    over = {k: {"scaled": round(v["max_scaled_err"], 3), "values_over": int(round(v["frac_over"] * v["n"])), "of": v["n"],
                "reference_f32_path_scaled": (round(v["reference_f32_noise"]["max_scaled_err"], 3)
                                              if "reference_f32_noise" in v else None)}
            for k, v in rep["checks"].items() if v["max_scaled_err"] > 1.0}
    worst5 = sorted(rep["checks"].items(), key=lambda kv: -kv[1]["max_scaled_err"])[:5]
    return {"max_scaled_err": rep["max_scaled_err"], "worst_check": rep["worst_check"], "n_checks": rep["n_checks"],
            "max_err_over_ref_max": rep["max_err_over_ref_max"], "worst_relative_check": rep["worst_relative_check"],
            "tolerance": rep["tolerance"], "sampled_rows": args.parity_rows, "checks_over_tolerance": over,
            "worst": {k: round(v["max_scaled_err"], 4) for k, v in worst5}, "seconds": round(time.time() - t0, 1),
            # the reference's OWN float32 arithmetic (sgemm + scipy float32) against the same float64 values, same
            # sample: the noise floor of a float32-vs-float32 comparison under this bound
            "reference_f32_noise": {k: {"scaled": round(v["max_scaled_err"], 3), "frac_over": v["frac_over"],
                                        "gpu_vs_reference_f32_scaled": round(v["gpu_vs_reference_f32"], 3)}
                                    for k, v in rep["reference_f32_noise"].items()},
            "max_scaled_err_over_reference_noise": rep["max_scaled_err_over_reference_noise"],
            "oracle": "scipy csr@dense (float32, the routine S.dot runs) on the GPU path's own operands per operation; float64 accumulation as the arbiter for dense contractions and N-long sums"}


def op_breakdown(m, dev):
    """One eager epoch with CUDA events around every libgcg call (warm, in situ): where the epoch goes."""
    import torch
    from graphconvgeo_b200 import ops
    names = ["spmm", "gemm", "colsum", "act_bwd", "highway_bwd", "softmax_ce", "sum_scaled", "scatter_rows"]
    orig = {n: getattr(ops, n) for n in names}
    rec = []

    def wrap(n, fn):
        def inner(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*a, **k)
            e.record()
            if n == "spmm":
                A, B = a[0], a[1]
                tag = "spmm %dx%d nnz=%d F=%d%s" % (A.shape[0], A.shape[1], getattr(A, "nnz", 0), B.shape[1],
                                                    " +gate" if k.get("gate") is not None else "")
            elif n == "gemm":
                A, B = a[0], a[1]
                tag = "gemm %s%s A%s B%s" % ("T" if k.get("transA") else "N", "T" if k.get("transB") else "N",
                                              tuple(A.shape), tuple(B.shape))
            else:
                tag = n + " " + "x".join(str(d) for d in a[0].shape)
            rec.append((tag, s, e))
            return r
        return inner

    g, m._graph = m._graph, None
    try:
        for n in names:
            setattr(ops, n, wrap(n, orig[n]))
        s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        m._train_step_enqueue()
        e0.record()
        torch.cuda.synchronize(dev)
    finally:
        for n in names:
            setattr(ops, n, orig[n])
        m._graph = g
    agg = {}
    for tag, s, e in rec:
        a = agg.setdefault(tag, [0, 0.0])
        a[0] += 1
        a[1] += s.elapsed_time(e)
    total = s0.elapsed_time(e0)
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
    if int(os.environ.get("RANK", "0")) == 0:
        log("---- op breakdown of one eager epoch: %.2f ms total ----" % total)
        for tag, (c, t) in rows:
            log("  %-70s x%-2d %8.3f ms  %5.1f%%" % (tag, c, t, 100 * t / total))
    return {"epoch_ms_eager": total, "ops": [{"op": tag, "calls": c, "ms": t} for tag, (c, t) in rows]}


def measure_e2e(m, args, dev, flush):
    """Epoch through the public call with HOST inputs: X (CSR), Y_train, train_indices are copied from pinned
    host memory every step (Theano copies f_train's inputs per call, mlpconv.py:295) and the step's [loss, acc]
    are read back.  The upload of step k+1 runs on a copy stream into one of two staging buffers while step k
    computes; a device-to-device copy (compute stream) then places it into the buffers the epoch reads.  The
    timed region starts before the first upload and holds all K uploads, K epochs and K read-backs."""
    import torch
    X = m.Xd
    devt = [X.indptr, X.indices, X.data, m.y_train_dev, m.ti_train.dev]
    host = [t.cpu().pin_memory() for t in devt]
    h2d = int(sum(t.numel() * t.element_size() for t in host))
    stage = [[torch.empty_like(t) for t in devt] for _ in range(2)]
    out_host = torch.empty(3, dtype=torch.float32).pin_memory()
    steps = max(3, min(args.steps, 10))
    main = torch.cuda.current_stream(dev)
    copy_stream = torch.cuda.Stream(dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def upload(k):
        b = k % 2
        copy_stream.wait_event(consumed[b])               # staging buffer b has been drained by step k - 2
        with torch.cuda.stream(copy_stream):
            for d, h in zip(stage[b], host):
                d.copy_(h, non_blocking=True)
            ready[b].record(copy_stream)

    def run(n_steps):
        for ev in consumed:
            ev.record(main)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(main)
        copy_stream.wait_event(s)
        upload(0)
        for k in range(n_steps):
            b = k % 2
            if flush is not None:
                flush.fill_(1)
            main.wait_event(ready[b])
            for d, st in zip(devt, stage[b]):
                d.copy_(st, non_blocking=True)
            consumed[b].record(main)
            if k + 1 < n_steps:
                upload(k + 1)
            hb = m.f_train()
            out_host[0:2].copy_(hb["out"], non_blocking=True)
            out_host[2:3].copy_(m.adam.reg_out, non_blocking=True)
        e.record(main)
        e.synchronize()
        return s.elapsed_time(e) / n_steps

    run(2)
    value = run(steps)
    return {"value": float(value), "unit": "ms", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
            "steps": steps, "upload": "pinned host -> staging on a copy stream (double buffered), overlapped with the "
                                      "previous epoch; device-to-device placement on the compute stream"}


def spmm_roofline(m, wl, dev, reps=10):
    """A_hat . H at F = hidden on the epoch's own operands, CUDA events around each launch."""
    import torch
    from graphconvgeo_b200 import ops
    peak, peak_src = measured_peaks()
    A = m.l_hid1.H
    H = m.l_hid1._out
    N, F = H.shape                      # local rows in row-partitioned mode
    out = ops.alloc_mat(N, F, dev)
    nnz = A.nnz
    n_in = A.shape[1]
    # compulsory traffic of this rank's launch(es): CSR once, the gathered operand once, the output once
    alg_bytes = 8 * nnz + 4 * (N + 1) + 4 * n_in * F + 4 * N * F
    if hasattr(A, "full"):      # feature-sliced: all rows, F/P columns on this rank (+ the two all-to-alls, timed too)
        fp = (-(-F // A.part.world) + 3) // 4 * 4
        alg_bytes = 8 * nnz + 4 * (n_in + 1) + 8 * n_in * fp
    results = {}
    dist_mode = hasattr(A, "dist_spmm")
    for label, panel in ((("auto", None),) if dist_mode else (("auto", None), ("rows", 0))):     # rows = register-gather kernel
        for _ in range(3):
            ops.spmm(A, H, out=out, panel_cols=panel)
        ts = []
        for _ in range(reps):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            ops.spmm(A, H, out=out, panel_cols=panel)
            e.record()
            e.synchronize()
            ts.append(s.elapsed_time(e))
        results[label] = float(np.mean(ts))
    t_ms = results["auto"]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tp) and not dist_mode:
        traffic = json.load(open(tp)).get(wl.name)
    others = None
    if not dist_mode:
        # the two other sparse products of the epoch against their own compulsory traffic
        from graphconvgeo_b200.lasagne_layers import _x_product, _xt_product
        l1 = m.l_hid1
        X = m.Xd
        V = X.shape[1]
        z = ops.alloc_mat(N, F, dev)
        dW = torch.empty_like(l1.W)

        def timed(fn):
            for _ in range(2):
                fn()
            ts = []
            for _ in range(5):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                fn()
                e.record()
                e.synchronize()
                ts.append(s.elapsed_time(e))
            return float(np.mean(ts))
        t_xw = timed(lambda: _x_product(l1, X, l1.W, z))
        t_xt = timed(lambda: _xt_product(l1, X, H, dW))
        b_xw = 8 * X.nnz + 4 * (N + 1) + 4 * V * F + 4 * N * F
        b_xt = 8 * X.nnz + 4 * (V + 1) + 4 * N * F + 4 * V * F
        others = [{"kernel": "X.W1 (a1): CSR [%d x %d] nnz %d times dense [%d x %d]%s" % (
                       N, V, X.nnz, V, F, (" (dense head of %d terms = %.0f%% of the non-zeros on tcgen05 + sparse tail)"
                                            % (l1._x_head[1].k_head, 100 * l1._x_head[1].head_fraction))
                       if getattr(l1, "_x_head", None) is not None else ""), "ms": t_xw,
                   "algorithmic_bytes": b_xw, "achieved": b_xw / t_xw / 1e6, "frac": b_xw / t_xw / 1e6 / peak, "unit": "GB/s",
                   "gathered_TBps": 4.0 * X.nnz * F / t_xw / 1e9},
                  {"kernel": "X^T.dZ1 (a6): document-blocked CSR of X^T [%d x %d] times dense [%d x %d]" % (V, N, N, F), "ms": t_xt,
                   "algorithmic_bytes": b_xt, "achieved": b_xt / t_xt / 1e6, "frac": b_xt / t_xt / 1e6 / peak, "unit": "GB/s",
                   "gathered_TBps": 4.0 * X.nnz * F / t_xt / 1e9}]
        del z, dW
    achieved = alg_bytes / (t_ms * 1e-3) / 1e9
    gather_model = (8 * nnz + 4 * nnz * F + 4 * N * F) / (t_ms * 1e-3) / 1e9
    return {"roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "%s (A_hat.H, F=%d%s)" % (
                             "spmm_stream_kernel" if ops.auto_panel_cols(
                                 n_in, ((-(-F // A.part.world) + 3) // 4 * 4) if hasattr(A, "full") else F, nnz) == -2
                             else "spmm_vec_kernel", F,
                             (", this rank's share: %s" % ("all rows x F/P columns incl. the two peer-store transposes"
                                                            if hasattr(A, "full") else "local rows incl. NCCL all-gather of H"))
                             if dist_mode else ""),
                         "algorithmic_bytes": alg_bytes, "ms": t_ms, "peak_source": peak_src,
                         "frac_of_nominal_8000": achieved / 8000.0, "other_sparse_products": others},
            "detail": {"N": N, "F": F, "nnz": nnz, "ms_auto": results["auto"], "ms_register_gather_kernel": results.get("rows"),
                       "panel_cols_auto": ops.auto_panel_cols(n_in, F, nnz), "gather_model_GBps": gather_model,
                       "plan": None if dist_mode else A.plan_info(),
                       "diag_fraction": getattr(A, "diag_fraction", None)}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="twitter-world", choices=["twitter-world", "twitter-us", "geotext", "tiny"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--highway", type=int, default=1)
    ap.add_argument("--random-graph", action="store_true", help="Chung-Lu graph without community structure")
    ap.add_argument("--no-graph", action="store_true",
                    help="time the epoch without a CUDA graph (SURVEY 8d config 2: the launch-bound GEOTEXT shape with and "
                         "without one); with GCG_NATIVE_EPOCH=1 the launches then come from gcg_epoch_run (C++), else from Python")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the sampled-row oracle check of one training step")
    ap.add_argument("--parity-rows", type=int, default=1024)
    ap.add_argument("--breakdown", action="store_true", help="per-op CUDA-event breakdown of one eager epoch")
    ap.add_argument("--no-peer-memory", action="store_true", help="feature mode: NCCL all-to-all instead of P2P stores")
    ap.add_argument("--partition", default="auto", choices=["auto", "feature", "row"],
                    help="multi-GPU scheme for A_hat.Z: feature slices + all-to-all, or row blocks + all-gather")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
