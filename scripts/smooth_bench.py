"""Measurement of SURVEY.md section 8(f) row 2 on one B200: the smoothing SpGEMM X_conv = A_hat * X
(main.py:528-530) and one epoch of the minibatch MLP on it (mlp.py:267-271), each next to the reference's
CPU path (scipy csr_matmat / the NumPy restatement in oracle/mlp_oracle.py) timed on the same box.

    python scripts/smooth_bench.py --workload geotext [--mlp] [--cpu-rows N]

Prints one JSON line (also written to gpurun_out/smooth_<workload>.json).  Not the headline bench: bench.py
keeps BASELINE.json's metric; this is the per-row measurement the scope table asks for.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from graphconvgeo_b200 import _lib, ops, synth  # noqa: E402
from graphconvgeo_b200.sparse import spgemm  # noqa: E402


def cuda_ms(fn, reps=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = None
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="geotext")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of A_hat the scipy baseline multiplies (0 = auto)")
    ap.add_argument("--mlp", action="store_true", help="also time one MLP epoch on the smoothed features")
    ap.add_argument("--hidden", type=int, default=500)      # main.py:496
    ap.add_argument("--batch", type=int, default=500)       # main.py:495
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    wl = synth.make_workload_device(args.workload, device=dev, seed=77, scale=args.scale)
    A, X = wl.A_hat, wl.X
    n, V = X.shape
    a64 = A.data.double()                                    # main.py:513-522 keeps A_hat in float64
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))           # fallback: B200_PROFILING.md 6.65 TB/s

    spgemm(A, X, a_values=a64)                               # warm-up (allocator, module load)
    ops.launch_count(reset=True)
    ms, Xc = cuda_ms(lambda: spgemm(A, X, a_values=a64))
    launches = ops.launch_count()
    deg_b = (X.indptr[1:] - X.indptr[:-1]).long()
    products = int(deg_b[A.indices.long()].sum().item())
    comp_bytes = 16 * A.nnz + 8 * X.nnz + 8 * Xc.nnz + 4 * (2 * n + 2) + 8 * (n + 1)   # A twice (2 passes, f64 vals once), B, C, offsets
    out = dict(metric="input smoothing SpGEMM X_conv = A_hat * X", workload=args.workload, n=n, vocab=V,
               nnz_A=A.nnz, nnz_X=X.nnz, nnz_out=Xc.nnz, products=products, gpu_ms=ms,
               gpu_products_per_s=products / (ms * 1e-3), gpu_launches=launches,
               roofline=dict(bound="hbm", achieved=comp_bytes / (ms * 1e-3) / 1e9, peak=hbm_peak, unit="GB/s",
                             frac=comp_bytes / (ms * 1e-3) / 1e9 / hbm_peak,
                             note="compulsory bytes (each array once per pass); the accumulator traffic "
                                  "(~products * 64 B of sectors) is what actually bounds the kernel"))

    # ---- reference CPU path: scipy csr_matmat on a bounded row sample, extrapolated by products
    import scipy.sparse as sp
    ip, ix = A.indptr.cpu().numpy(), A.indices.cpu().numpy()
    Xs = X.to_scipy()
    rows = args.cpu_rows or max(1, min(n, int(n * min(1.0, 1.5e8 / max(products, 1)))))
    Asub = sp.csr_matrix((a64.cpu().numpy()[:ip[rows]], ix[:ip[rows]], ip[:rows + 1]), shape=(rows, n))
    t0 = time.perf_counter()
    ref = (Asub * Xs).tocsr().astype("float32")
    cpu_s = time.perf_counter() - t0
    sub_products = int(np.diff(Xs.indptr)[ix[:ip[rows]]].sum())
    cpu_full_s = cpu_s * products / max(sub_products, 1)
    # parity of the sample while we are here (bit-exact)
    gi = Xc.indptr[:rows + 1].cpu().numpy()
    same = (np.array_equal(gi, ref.indptr) and np.array_equal(Xc.indices[:gi[-1]].cpu().numpy(), ref.indices)
            and np.array_equal(Xc.data[:gi[-1]].cpu().numpy(), ref.data))
    out["cpu_baseline"] = dict(kind="reference (scipy csr_matmat, the routine H * X calls)", cores=1,
                               sample="first %d of %d rows, extrapolated by product count" % (rows, n),
                               sample_s=cpu_s, value=cpu_full_s * 1e3, unit="ms")
    out["bit_exact_on_sample"] = bool(same)
    out["speedup_vs_cpu"] = cpu_full_s * 1e3 / ms

    if args.mlp:
        from graphconvgeo_b200.mlp import MLP
        ntr = len(wl.train_indices)
        Xc_host = Xc.to_scipy()
        y = wl.Y.astype(np.int32)
        uniq = np.unique(y[:ntr])
        remap = -np.ones(int(y.max()) + 1, np.int64)
        remap[uniq] = np.arange(len(uniq))
        ytr = remap[y[:ntr]].astype(np.int32)
        clf = MLP(n_epochs=1, batch_size=args.batch, regul_coefs=[1e-6, 1e-6], hidden_layer_size=args.hidden,
                  drop_out=False, seed=0)
        clf.prepare(Xc_host[:ntr], ytr)
        clf.train_epoch()                                    # warm-up
        ops.launch_count(reset=True)
        if os.environ.get("GCG_PROFILE_EPOCH"):
            import cProfile
            import pstats
            pr = cProfile.Profile()
            pr.enable()
            clf.train_epoch()
            torch.cuda.synchronize()
            pr.disable()
            pstats.Stats(pr, stream=sys.stderr).sort_stats("tottime").print_stats(14)
        ems, nb = cuda_ms(clf.train_epoch)
        launches_epoch = ops.launch_count()
        all_ms = [ems] + [cuda_ms(clf.train_epoch)[0] for _ in range(4)]
        ems = float(np.median(all_ms))
        mlp = dict(n_train=ntr, batch=args.batch, hidden=args.hidden, classes=len(uniq), batches_per_epoch=nb,
                   gpu_epoch_ms=ems, gpu_epoch_ms_all=all_ms, gpu_launches=launches_epoch)
        import bench                                         # its cpu_baseline leg is the only oracle user
        cpu_ms, sample_batches = bench.cpu_mlp_epoch(Xc_host[:ntr], ytr, args.hidden, len(uniq), args.batch, nb)
        mlp["cpu_baseline"] = dict(kind="port", cores=1, sample="%d minibatches" % sample_batches,
                                   value=cpu_ms, unit="ms per epoch")
        mlp["speedup_vs_cpu"] = cpu_ms / ems
        out["mlp_epoch"] = mlp

    line = json.dumps(out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "smooth_%s.json" % args.workload), "w") as f:
        f.write(line + "\n")
    print(line)


if __name__ == "__main__":
    if not torch.cuda.is_available():
        raise SystemExit("smooth_bench.py needs a CUDA device (no CPU fallback)")
    _lib.lib()
    main()
