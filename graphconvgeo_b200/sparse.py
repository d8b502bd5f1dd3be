"""Device CSR matrices and their SpMM plans.

The reference hands scipy CSR matrices (X: data.py:387-394, A_hat:
tensormain.py:168-181) to Theano, which wraps them as sparse variables
(mlpconv.py:178-180).  Here a ``CSRMatrix`` holds the three CSR arrays as torch
CUDA tensors (device memory plumbing only) plus the opaque ``gcg_plan``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _np_ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def is_sparse(x) -> bool:
    """The type test of lasagne_layers.py:22-24 / :33-35 / :61-63."""
    if isinstance(x, CSRMatrix):
        return True
    try:
        import scipy.sparse as sp
        return sp.issparse(x)
    except Exception:       # pragma: no cover
        return False


class CSRMatrix:
    """float32 CSR on one GPU.  ``host`` keeps the (indptr, indices, data) NumPy
    arrays when the matrix came from the host so that transposes / row gathers
    (one-off, per fit) can be done by the C++ host helpers."""

    def __init__(self, indptr, indices, data, shape, host=None, long_row_threshold=256):
        self.indptr, self.indices, self.data = indptr, indices, data
        self.shape = (int(shape[0]), int(shape[1]))
        self.host = host
        self.long_row_threshold = int(long_row_threshold)
        self._plan = None
        self._T = None
        assert indptr.dtype == torch.int32 and indices.dtype == torch.int32 and data.dtype == torch.float32
        assert indptr.numel() == self.shape[0] + 1

    # ------------------------------------------------------------------ ctor
    @classmethod
    def from_scipy(cls, m, device="cuda", long_row_threshold=256):
        import scipy.sparse as sp
        m = sp.csr_matrix(m)
        if not m.has_sorted_indices:
            m = m.copy()
            m.sort_indices()
        if m.nnz >= 2 ** 31 - 1:
            raise ValueError("nnz does not fit int32 CSR indices")
        host = (np.ascontiguousarray(m.indptr, dtype=np.int32),
                np.ascontiguousarray(m.indices, dtype=np.int32),
                np.ascontiguousarray(m.data, dtype=np.float32))
        return cls.from_host(host, m.shape, device, long_row_threshold)

    @classmethod
    def from_host(cls, host, shape, device="cuda", long_row_threshold=256):
        dev = torch.device(device)
        t = [torch.from_numpy(a).to(dev, non_blocking=False) for a in host]
        return cls(t[0], t[1], t[2], shape, host=host, long_row_threshold=long_row_threshold)

    # ------------------------------------------------------------------ info
    @property
    def nnz(self):
        return int(self.indices.numel())

    @property
    def device(self):
        return self.data.device

    def astype(self, dtype):
        if np.dtype(dtype) != np.float32:
            raise ValueError("the B200 hot path is float32 only (mlpconv.py:136)")
        return self

    def _host_arrays(self):
        if self.host is None:
            self.host = (self.indptr.cpu().numpy(), self.indices.cpu().numpy(), self.data.cpu().numpy())
        return self.host

    def to_scipy(self):
        import scipy.sparse as sp
        ip, ix, d = self._host_arrays()
        return sp.csr_matrix((d, ix, ip), shape=self.shape)

    # ------------------------------------------------------------------ plan
    @property
    def plan(self):
        if self._plan is None:
            L = _lib.lib()
            out = C.c_void_p()
            hp = self.host[0] if self.host is not None else None
            _lib.check(L.gcg_plan_create_csr(self.shape[0], self.shape[1], self.nnz,
                                             self.indptr.data_ptr(), self.indices.data_ptr(),
                                             self.data.data_ptr(), _np_ptr(hp) if hp is not None else None,
                                             self.long_row_threshold, C.byref(out)), "gcg_plan_create_csr")
            self._plan = out
        return self._plan

    def plan_info(self):
        info = (C.c_int64 * 8)()
        _lib.check(_lib.lib().gcg_plan_info(self.plan, info), "gcg_plan_info")
        keys = ["n_rows", "n_cols", "nnz", "n_long_rows", "n_segments", "max_degree", "long_row_threshold"]
        return dict(zip(keys, list(info)[:7]))

    def workspace_bytes(self, ldc):
        return int(_lib.lib().gcg_plan_workspace_bytes(self.plan, int(ldc)))

    def __del__(self):
        try:
            if self._plan is not None:
                _lib.lib().gcg_plan_destroy(self._plan)
                self._plan = None
        except Exception:
            pass

    # ------------------------------------------------- one-off host helpers
    @property
    def T(self):
        """CSR of the transpose (stable), built once on the host (gcg_csr_transpose_host)."""
        if self._T is None:
            ip, ix, d = self._host_arrays()
            n, m = self.shape
            tip = np.empty(m + 1, np.int32)
            tix = np.empty(len(ix), np.int32)
            td = np.empty(len(ix), np.float32)
            _lib.check(_lib.lib().gcg_csr_transpose_host(n, m, _np_ptr(ip), _np_ptr(ix), _np_ptr(d),
                                                         _np_ptr(tip), _np_ptr(tix), _np_ptr(td)),
                       "gcg_csr_transpose_host")
            self._T = CSRMatrix.from_host((tip, tix, td), (m, n), self.device, self.long_row_threshold)
        return self._T

    def gather_rows(self, idx):
        """CSR of rows ``idx`` (duplicates allowed): A[idx, :] (gcg_csr_gather_rows_host)."""
        idx = np.ascontiguousarray(np.asarray(idx), dtype=np.int32)
        ip, ix, d = self._host_arrays()
        L = _lib.lib()
        nnz = L.gcg_csr_gather_rows_host(self.shape[0], _np_ptr(ip), _np_ptr(ix), _np_ptr(d), _np_ptr(idx),
                                         len(idx), None, None, None)
        if nnz < 0:
            _lib.check(-1, "gcg_csr_gather_rows_host")
        oip = np.empty(len(idx) + 1, np.int32)
        oix = np.empty(nnz, np.int32)
        od = np.empty(nnz, np.float32)
        r = L.gcg_csr_gather_rows_host(self.shape[0], _np_ptr(ip), _np_ptr(ix), _np_ptr(d), _np_ptr(idx),
                                       len(idx), _np_ptr(oip), _np_ptr(oix), _np_ptr(od))
        if r < 0:
            _lib.check(-1, "gcg_csr_gather_rows_host")
        return CSRMatrix.from_host((oip, oix, od), (len(idx), self.shape[1]), self.device,
                                   self.long_row_threshold)


    def permute(self, order, col_map=None):
        """P A Q^T: row i of the result is row order[i]; column j becomes col_map[j]
        (None keeps the columns).  One-off host operation (gcg_csr_permute_host)."""
        order = np.ascontiguousarray(np.asarray(order), dtype=np.int32)
        cm = None if col_map is None else np.ascontiguousarray(np.asarray(col_map), dtype=np.int32)
        ip, ix, d = self._host_arrays()
        n = self.shape[0]
        assert len(order) == n
        oip = np.empty(n + 1, np.int32)
        oix = np.empty(len(ix), np.int32)
        od = np.empty(len(ix), np.float32)
        _lib.check(_lib.lib().gcg_csr_permute_host(n, _np_ptr(ip), _np_ptr(ix), _np_ptr(d), _np_ptr(order),
                                                   _np_ptr(cm) if cm is not None else None, _np_ptr(oip),
                                                   _np_ptr(oix), _np_ptr(od)), "gcg_csr_permute_host")
        return CSRMatrix.from_host((oip, oix, od), self.shape, self.device, self.long_row_threshold)


def as_csr(x, device="cuda", long_row_threshold=256) -> CSRMatrix:
    if isinstance(x, CSRMatrix) or hasattr(x, "dist_spmm"):      # device CSR or a row-partitioned one
        return x
    return CSRMatrix.from_scipy(x, device=device, long_row_threshold=long_row_threshold)


# --------------------------------------------------------------------------- #
# host-side graph preparation                                                  #
# --------------------------------------------------------------------------- #


def build_ahat_host(adj):
    """A_hat = D^-1/2 (A, unit diagonal) D^-1/2, float64 then float32 -- the
    normalisation of tensormain.py:170-180,221, done by gcg_ahat_build_host.
    ``adj``: scipy sparse adjacency (binary or weighted).  Returns scipy CSR float32."""
    import scipy.sparse as sp
    a = sp.csr_matrix(adj)
    if not a.has_sorted_indices:
        a = a.copy()
        a.sort_indices()
    a.sum_duplicates()
    n = a.shape[0]
    assert a.shape[0] == a.shape[1]
    ip = np.ascontiguousarray(a.indptr, dtype=np.int32)
    ix = np.ascontiguousarray(a.indices, dtype=np.int32)
    binary = bool(np.all(a.data == 1))
    w = None if binary else np.ascontiguousarray(a.data, dtype=np.float64)
    L = _lib.lib()
    nnz = L.gcg_ahat_nnz_host(n, _np_ptr(ip), _np_ptr(ix))
    if nnz < 0:
        raise _lib.GcgError("gcg_ahat_nnz_host failed")
    oip = np.empty(n + 1, np.int32)
    oix = np.empty(nnz, np.int32)
    ov = np.empty(nnz, np.float32)
    _lib.check(L.gcg_ahat_build_host(n, _np_ptr(ip), _np_ptr(ix), _np_ptr(w) if w is not None else None,
                                     _np_ptr(oip), _np_ptr(oix), _np_ptr(ov)), "gcg_ahat_build_host")
    return sp.csr_matrix((ov, oix, oip), shape=(n, n))


def build_ahat_device(indptr, indices, n):
    """Device-side A_hat (tensormain.py:170-180,221) from a binary adjacency pattern already on the GPU
    (int32 CSR, sorted columns).  Bit-identical to build_ahat_host; returns a CSRMatrix."""
    dev = indptr.device
    L = _lib.lib()
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    wsb = int(L.gcg_ahat_device_workspace_bytes(n))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    out_ip = torch.empty(n + 1, dtype=torch.int32, device=dev)
    total = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(L.gcg_ahat_indptr_device(n, indptr.data_ptr(), indices.data_ptr(), out_ip.data_ptr(), total.data_ptr(),
                                        ws.data_ptr(), wsb, stream), "gcg_ahat_indptr_device")
    nnz = int(indices.numel()) + int(total.item())
    out_ix = torch.empty(nnz, dtype=torch.int32, device=dev)
    out_v = torch.empty(nnz, dtype=torch.float32, device=dev)
    _lib.check(L.gcg_ahat_fill_device(n, indptr.data_ptr(), indices.data_ptr(), out_ip.data_ptr(), out_ix.data_ptr(),
                                      out_v.data_ptr(), ws.data_ptr(), wsb, stream), "gcg_ahat_fill_device")
    return CSRMatrix(out_ip, out_ix, out_v, (n, n))
