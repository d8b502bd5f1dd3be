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
def cpu_reference_epoch(workload, n_layers, highway, scale, epochs=1, threads=None, warmup=0):
    """The reference's CPU path (oracle port: scipy csr@dense + BLAS, float32) on a bounded sample
    of the workload: the same shape generator at `scale` of the node/vocabulary counts.  Returns
    (mean ms per epoch on the sample over `epochs` timed epochs, description, threads)."""
    from graphconvgeo_b200 import synth
    from oracle import gcn_oracle as go
    threads = threads or os.cpu_count()
    w = synth.make_workload(workload, scale=scale)
    rng = np.random.RandomState(0)
    params = go.init_params(rng, w.X.shape[1], w.hidden, w.n_classes, n_layers, highway)
    net = go.GCNOracle(w.X, w.A_hat, n_layers, highway, (1e-6, 1e-6))
    y = w.Y[w.train_indices].astype(np.int32)
    st = go.AdamState(params)
    times = []
    for i in range(warmup + epochs):
        t0 = time.perf_counter()
        loss, acc, grads, _c = net.loss_and_grads(params, w.train_indices, y)
        go.adam_step(params, grads, st)
        dt = (time.perf_counter() - t0) * 1e3
        if i >= warmup:
            times.append(dt)
        if sum(times) > 120e3:
            break
    desc = ("%d epoch(s) of the oracle port (scipy csr@dense, 1 thread as under Theano, + BLAS on %d threads) on the %s "
            "generator at scale %.4g: %d nodes, vocab %d, nnzA %d, nnzX %d; value = sample ms / scale"
            % (len(times), threads, workload, scale, w.meta["n"], w.X.shape[1], w.meta["nnz_A"], w.meta["nnz_X"]))
    return float(np.mean(times)), desc, threads, len(times)


CPU_SCALE = {"twitter-world": 1.0 / 64, "twitter-us": 1.0 / 24, "geotext": 1.0, "tiny": 1.0}


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
    installable here) on the box's host cores, bounded sample, same metric/unit/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scale = CPU_SCALE[args.workload]
    ms_sample, desc, threads, n_timed = cpu_reference_epoch(args.workload, args.layers, bool(args.highway), scale,
                                                            epochs=args.steps, warmup=min(args.warmup, 1))
    ms = ms_sample / scale
    line = {
        "impl": "reference", "metric": "gcn_fwd_bwd_epoch_ms", "value": ms, "unit": "ms", "n_gpus": args.gpus,
        "steps": n_timed, "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args),
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
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    t0 = time.time()
    wl = synth.make_workload_device(args.workload, device=dev, seed=77, community=not args.random_graph)
    log("[rank %d] workload %s built in %.1fs: %s" % (rank, args.workload, time.time() - t0, wl.meta))

    model_kwargs = dict(n_epochs=args.steps, regul_coefs=[1e-6, 1e-6], hidden_layer_size=wl.hidden, drop_out=False,
                        n_layers=args.layers, highway=bool(args.highway), seed=1, device=dev, cuda_graph=True)
    if world > 1:
        from graphconvgeo_b200.dist import DistMLPCONV
        m = DistMLPCONV(partition=args.partition, peer_memory=not args.no_peer_memory, **model_kwargs)
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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        scale = CPU_SCALE[args.workload]
        ms, desc, threads, _n = cpu_reference_epoch(args.workload, args.layers, bool(args.highway), scale, epochs=1)
        cpu = {"value": ms / scale, "unit": "ms", "cores": threads, "kind": "port", "sample": desc,
               "sample_ms": ms}

    line = {
        "metric": "gcn_fwd_bwd_epoch_ms", "value": ms_per_step, "unit": "ms", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(bench_config(args), nodes=wl.meta["n"], vocab=wl.meta["vocab"], hidden=wl.hidden,
                       regions=wl.n_classes, nnz_A=wl.meta["nnz_A"], nnz_X=wl.meta["nnz_X"],
                       max_degree=wl.meta["max_degree"], graph="community" if wl.meta["community"] else "chung-lu",
                       parallelism=("rows x%d, A_hat.Z %s" % (world, ("feature-sliced, transposes by %s" % (
                           "peer-memory stores over NVLink" if getattr(getattr(m, "part", None), "peer", None) is not None
                           else "NCCL all-to-all")) if getattr(m, "partition", "") == "feature" else "row blocks + NCCL all-gather"))
                       if world > 1 else "single GPU",
                       gemm_mode=os.environ.get("GCG_GEMM_MODE", "auto")),
        "clocks": sampler.summary(), "gpu_launches": launches * args.steps,
        "launches_per_epoch": launches, "loss": loss, "acc": acc,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if roof is not None:
        line["roofline"] = roof["roofline"]
        line["spmm"] = roof["detail"]
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if breakdown is not None:
        line["breakdown"] = breakdown
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


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
    """Epoch through the public call with HOST inputs: X (CSR), Y_train, train_indices are copied from
    pinned host memory every step (Theano copies f_train's inputs per call, mlpconv.py:295) and the
    step's [loss, acc] are read back."""
    import torch
    X = m.Xd
    host = [t.cpu().pin_memory() for t in (X.indptr, X.indices, X.data, m.y_train_dev, m.ti_train.dev)]
    devt = [X.indptr, X.indices, X.data, m.y_train_dev, m.ti_train.dev]
    h2d = int(sum(t.numel() * t.element_size() for t in host))
    out_host = torch.empty(3, dtype=torch.float32).pin_memory()
    steps = max(3, min(args.steps, 10))
    times = []
    for i in range(steps + 2):
        if flush is not None:
            flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for d, h in zip(devt, host):
            d.copy_(h, non_blocking=True)
        hb = m.f_train()
        out_host[0:2].copy_(hb["out"], non_blocking=True)
        out_host[2:3].copy_(m.adam.reg_out, non_blocking=True)
        e.record()
        e.synchronize()
        if i >= 2:
            times.append(s.elapsed_time(e))
    return {"value": float(np.mean(times)), "unit": "ms", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
            "steps": steps}


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
    for label, panel in ((("auto", None),) if dist_mode else (("auto", None), ("rows", 0))):
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
    tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tp) and not dist_mode:
        traffic = json.load(open(tp)).get(wl.name)
    achieved = alg_bytes / (t_ms * 1e-3) / 1e9
    gather_model = (8 * nnz + 4 * nnz * F + 4 * N * F) / (t_ms * 1e-3) / 1e9
    return {"roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "spmm_vec_kernel (A_hat.H, F=%d%s)" % (F, ", local rows incl. NCCL all-gather of H" if dist_mode else ""),
                         "algorithmic_bytes": alg_bytes, "ms": t_ms, "peak_source": peak_src,
                         "frac_of_nominal_8000": achieved / 8000.0},
            "detail": {"N": N, "F": F, "nnz": nnz, "ms_auto_panel": results["auto"], "ms_whole_rows": results.get("rows"),
                       "panel_cols_auto": ops.auto_panel_cols(N, F), "gather_model_GBps": gather_model,
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
