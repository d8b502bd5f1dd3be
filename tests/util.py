import numpy as np

# north_star tolerance for float32 activations / gradients: |a-b| <= 1e-6 + 1e-4*|b|
ATOL, RTOL = 1e-6, 1e-4


def assert_close(got, ref, atol=ATOL, rtol=RTOL, what=""):
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (what, got.shape, ref.shape)
    err = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    tol = atol + rtol * np.abs(ref.astype(np.float64))
    bad = err > tol
    if bad.any():
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError("%s: %d/%d elements out of tolerance; worst at %s: got %r ref %r (err %.3e, tol %.3e)"
                             % (what, int(bad.sum()), bad.size, i, got[i], ref[i], err[i], tol[i]))


def to_dev(a, dev="cuda"):
    import torch
    from graphconvgeo_b200 import ops
    a = np.asarray(a, dtype=np.float32)
    if a.ndim == 2:
        out = ops.alloc_mat(a.shape[0], a.shape[1], dev)
        out.copy_(torch.from_numpy(np.ascontiguousarray(a)))
        return out
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def random_csr(rng, n_rows, n_cols, avg_deg, hub_rows=(), hub_deg=0, empty_frac=0.1, dtype=np.float32):
    """Random CSR with empty rows and optional hub rows (power-law stand-in)."""
    import scipy.sparse as sp
    deg = rng.poisson(avg_deg, size=n_rows)
    deg[rng.rand(n_rows) < empty_frac] = 0
    for r in hub_rows:
        deg[r] = hub_deg
    deg = np.minimum(deg, n_cols)
    indptr = np.zeros(n_rows + 1, np.int64)
    np.cumsum(deg, out=indptr[1:])
    indices = np.empty(indptr[-1], np.int64)
    for r in range(n_rows):
        indices[indptr[r]:indptr[r + 1]] = np.sort(rng.choice(n_cols, size=deg[r], replace=False))
    data = rng.standard_normal(indptr[-1]).astype(dtype)
    return sp.csr_matrix((data, indices, indptr), shape=(n_rows, n_cols))
