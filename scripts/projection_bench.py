"""Measurement of SURVEY.md section 8(f) row 4 on one B200: celebrity filter + projection of a synthetic
@-mention graph (data.py:364-373) as the pattern SpGEMM R^T R, next to the plain-Python port of the
reference loop (bounded sample) and scipy's boolean product, all on the same box.

    python scripts/projection_bench.py [--targets 450000] [--names 900000] [--mentions 8]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from graphconvgeo_b200 import _lib, ops  # noqa: E402
from graphconvgeo_b200.graph import mention_incidence, remove_celebrities  # noqa: E402
from graphconvgeo_b200.sparse import CSRMatrix, spgemm_pattern  # noqa: E402


def mention_graph(n_targets, n_names, mentions, p_target, seed=77):
    rng = np.random.RandomState(seed)
    k = rng.poisson(mentions, size=n_targets)
    users = np.repeat(np.arange(n_targets), k)
    tot = len(users)
    to_target = rng.rand(tot) < p_target
    names = n_targets + np.minimum(rng.zipf(1.3, size=tot) - 1, n_names - 1)
    dst = np.where(to_target, rng.randint(0, n_targets, size=tot), names)
    M = n_targets + n_names
    B = sp.coo_matrix((np.ones(tot, np.float32), (users, dst)), shape=(M, M)).tocsr()
    B = B + B.T + sp.eye(M, format="csr", dtype=np.float32).multiply(
        sp.csr_matrix((np.ones(n_targets, np.float32), (np.arange(n_targets), np.arange(n_targets))), shape=(M, M)))
    B = sp.csr_matrix(B)
    B.data[:] = 1.0
    return B


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--targets", type=int, default=450000)
    ap.add_argument("--names", type=int, default=900000)
    ap.add_argument("--mentions", type=float, default=8.0)
    ap.add_argument("--p-target", type=float, default=0.2)
    ap.add_argument("--threshold", type=int, default=10)      # DataLoader default, data.py:255
    args = ap.parse_args()
    n = args.targets
    B = mention_graph(n, args.names, args.mentions, args.p_target)
    t0 = time.perf_counter()
    Bf = remove_celebrities(B, n, args.threshold)
    R = mention_incidence(Bf, n)
    RT = sp.csr_matrix(R.T)
    RT.sort_indices()
    host_prep_s = time.perf_counter() - t0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    Rd, RTd = CSRMatrix.from_scipy(R), CSRMatrix.from_scipy(RT)
    G = spgemm_pattern(RTd, Rd, drop_diagonal=True)
    torch.cuda.synchronize()
    first_s = time.perf_counter() - t0
    ops.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    G = spgemm_pattern(RTd, Rd, drop_diagonal=True)
    e1.record()
    torch.cuda.synchronize()
    gpu_ms = e0.elapsed_time(e1)
    launches = ops.launch_count()
    products = int(np.diff(R.indptr)[RT.indices].sum())
    t0 = time.perf_counter()
    Gs = sp.csr_matrix(RT @ R)
    Gs.setdiag(0)
    Gs.eliminate_zeros()
    Gs.sort_indices()
    scipy_ms = (time.perf_counter() - t0) * 1e3
    same = np.array_equal(G.indptr.cpu().numpy(), Gs.indptr) and np.array_equal(G.indices.cpu().numpy(), Gs.indices)
    import bench                                             # its cpu_baseline leg is the only oracle user
    port_ms, sampled = bench.cpu_projection(Bf, n)
    out = dict(metric="mention-graph projection (celebrity filter + R^T R pattern)", targets=n,
               nodes=int(B.shape[0]), mention_edges=int(B.nnz // 2), kept_incidence_nnz=int(R.nnz),
               projected_edges=int(G.nnz // 2), products=products, gpu_ms=gpu_ms, gpu_launches=launches,
               gpu_products_per_s=products / (gpu_ms * 1e-3), first_call_with_upload_ms=first_s * 1e3,
               host_prep_ms=host_prep_s * 1e3, identical_to_scipy_pattern=bool(same),
               cpu_baseline=dict(kind="port", cores=1, value=port_ms, unit="ms",
                                 sample="plain-Python data.py:226-250 loop over the first %d of %d nodes, scaled by "
                                        "clique-pair count" % (sampled, B.shape[0])),
               scipy_boolean_product_ms=scipy_ms, speedup_vs_port=port_ms / gpu_ms,
               speedup_vs_scipy=scipy_ms / gpu_ms)
    line = json.dumps(out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "projection.json"), "w").write(line + "\n")
    print(line)


if __name__ == "__main__":
    if not torch.cuda.is_available():
        raise SystemExit("projection_bench.py needs a CUDA device (no CPU fallback)")
    _lib.lib()
    main()
