"""Device CSR matrices and their SpMM plans.

The reference hands scipy CSR matrices (X: data.py:387-394, A_hat:
tensormain.py:168-181) to Theano, which wraps them as sparse variables
(mlpconv.py:178-180).  Here a ``CSRMatrix`` holds the three CSR arrays as torch
CUDA tensors (device memory plumbing only) plus the opaque ``gcg_plan``.
"""
from __future__ import annotations

import ctypes as C
import os as _os

import numpy as np
import torch

from . import _lib


def _np_ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def is_sparse(x) -> bool:
    """The type test of lasagne_layers.py:22-24 / :33-35 / :61-63."""
    if isinstance(x, (CSRMatrix, RowBlockedCSR)):
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
        self.device_built = False     # True: no host copy of indices/data exists (device-sliced minibatch)
        assert indptr.dtype == torch.int32 and indices.dtype == torch.int32 and data.dtype == torch.float32
        assert indptr.numel() == self.shape[0] + 1

    # ------------------------------------------------------------------ ctor
    @classmethod
    def from_scipy(cls, m, device="cuda", long_row_threshold=256, sort_indices=True):
        import scipy.sparse as sp
        m = sp.csr_matrix(m)
        if sort_indices and not m.has_sorted_indices:
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
        if self.host is None or self.host[1] is None:
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

    def set_schedule(self, block_rows=None, block_panels=None):
        """Row-block x column-panel schedule of the streaming SpMM (gcg_plan_set_schedule).  ``block_rows``:
        ascending row offsets [0, ..., n_rows]; ``block_panels``: panels per block (> 0 panel-major inside the
        block, < 0 interleaved).  ``None`` restores whole rows in one block."""
        if block_rows is None:
            _lib.check(_lib.lib().gcg_plan_set_schedule(self.plan, 0, None, None), "gcg_plan_set_schedule")
            self.schedule = None
            return
        br = np.ascontiguousarray(np.asarray(block_rows), dtype=np.int32)
        bp = np.ascontiguousarray(np.asarray(block_panels), dtype=np.int32)
        if len(br) != len(bp) + 1:
            raise ValueError("set_schedule: need len(block_rows) == len(block_panels) + 1")
        _lib.check(_lib.lib().gcg_plan_set_schedule(self.plan, len(bp), _np_ptr(br), _np_ptr(bp)),
                   "gcg_plan_set_schedule")
        self.schedule = (br, bp)

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
        """CSR of the transpose (stable), built once on the host (gcg_csr_transpose_host); matrices
        sliced on the device (minibatches) are transposed there (gcg_csr_transpose_device)."""
        if self._T is None and self.device_built:
            return self.transpose_device()
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

    # ------------------------------------------------ device-side minibatch helpers
    def gather_rows_device(self, rows):
        """A[rows, :] built on the GPU (gcg_csr_gather_rows_device) -- `inputs[excerpt]` of
        iterate_minibatches (mlp.py:81-91).  ``rows``: host int array; entry order inside rows is kept."""
        rows = np.ascontiguousarray(np.asarray(rows), dtype=np.int32)
        if len(rows) and (rows.min() < 0 or rows.max() >= self.shape[0]):
            raise IndexError("row index out of range for a matrix with %d rows" % self.shape[0])
        if self.host is None:
            self.host = (self.indptr.cpu().numpy(), None, None)
        ip = self.host[0]
        lens = (ip[1:] - ip[:-1])[rows]
        oip = np.zeros(len(rows) + 1, np.int64)
        np.cumsum(lens, out=oip[1:])
        nnz = int(oip[-1])
        if nnz >= 2 ** 31 - 1:
            raise ValueError("gathered rows do not fit int32 CSR offsets")
        oip = oip.astype(np.int32)
        dev = self.device
        d_rows = torch.from_numpy(rows).to(dev)
        d_oip = torch.from_numpy(oip).to(dev)
        oix = torch.empty(nnz, dtype=torch.int32, device=dev)
        od = torch.empty(nnz, dtype=torch.float32, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(_lib.lib().gcg_csr_gather_rows_device(self.indptr.data_ptr(), self.indices.data_ptr(),
                                                         self.data.data_ptr(), d_rows.data_ptr(), len(rows),
                                                         d_oip.data_ptr(), oix.data_ptr(), od.data_ptr(), stream),
                   "gcg_csr_gather_rows_device")
        out = CSRMatrix(d_oip, oix, od, (len(rows), self.shape[1]), host=(oip, None, None),
                        long_row_threshold=self.long_row_threshold)
        out.device_built = True           # no host copy of indices/data: transposes happen on the device too
        return out

    def transpose_device(self):
        """Stable CSR transpose on the GPU (gcg_csr_transpose_device); cached like ``T``."""
        if self._T is None:
            n, m = self.shape
            dev = self.device
            L = _lib.lib()
            nnz = self.nnz
            wsb = int(L.gcg_csr_transpose_device_workspace_bytes(n, m, nnz))
            ws = torch.empty(wsb + 256, dtype=torch.uint8, device=dev)
            off = (-ws.data_ptr()) % 256
            tip = torch.empty(m + 1, dtype=torch.int32, device=dev)
            tix = torch.empty(nnz, dtype=torch.int32, device=dev)
            td = torch.empty(nnz, dtype=torch.float32, device=dev)
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(L.gcg_csr_transpose_device(n, m, nnz, self.indptr.data_ptr(), self.indices.data_ptr(),
                                                  self.data.data_ptr(), tip.data_ptr(), tix.data_ptr(), td.data_ptr(),
                                                  ws.data_ptr() + off, wsb, stream), "gcg_csr_transpose_device")
            self._T = CSRMatrix(tip, tix, td, (m, n), long_row_threshold=self.long_row_threshold)
        return self._T

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


def l2_schedule(A: "CSRMatrix", F, budget_bytes=48 << 20, min_panel_floats=128,
                block_rows_candidates=(16384, 65536, 262144), report=None):
    """Row-block x column-panel schedule for the streaming SpMM of ``A`` against an [n_cols, F] operand
    (CSRMatrix.set_schedule), chosen by a traffic model evaluated on the matrix itself.

    For a block of consecutive rows, D = distinct columns it references.  Split into p column panels, the
    block gathers from D * 4F/p bytes per panel; when that fits ``budget_bytes`` of L2 every repeated reference
    hits, otherwise only budget / working-set of them.  Per candidate block height the model sums first
    touches, missed repeats, the CSR re-read per panel and a fixed per-non-zero cost per extra panel pass,
    and the cheapest candidate wins.  Index analysis (sort / unique of nnz keys) runs with torch on the
    matrix's device -- plumbing, like the rest of the one-off plan preparation."""
    n, ncols = A.shape
    dev = A.device
    row_bytes = 4.0 * F
    deg = (A.indptr[1:] - A.indptr[:-1]).to(torch.int64)
    rows = torch.repeat_interleave(torch.arange(n, device=dev, dtype=torch.int64), deg)
    cols = A.indices.to(torch.int64)
    max_panels = max(1, int(F) // int(min_panel_floats))
    best = None
    for R in block_rows_candidates:
        nb = -(-n // R)
        blk = rows // R
        uniq = torch.unique(blk * ncols + cols)
        D = torch.bincount(uniq // ncols, minlength=nb).double()
        del uniq
        nnz_b = torch.bincount(blk, minlength=nb).double()
        panels = torch.clamp(torch.ceil(D * row_bytes / budget_bytes), 1, max_panels)
        ws = D * row_bytes / panels
        hit = torch.clamp(budget_bytes / torch.clamp(ws, min=1.0), max=1.0)
        est = (D * row_bytes + (nnz_b - D) * row_bytes * (1.0 - hit) + nnz_b * 8.0 * panels
               + nnz_b * 256.0 * (panels - 1.0)).sum().item()
        cand = dict(block_rows=R, est_bytes=est, panels=panels.to(torch.int32).cpu().numpy(),
                    distinct=float(D.sum().item()))
        if report is not None:
            report.append({k: (v if k != "panels" else np.bincount(v).tolist()) for k, v in cand.items()})
        if best is None or est < best["est_bytes"]:
            best = cand
        if nb == 1:
            break
    R = best["block_rows"]
    nb = -(-n // R)
    br = np.minimum(np.arange(nb + 1, dtype=np.int64) * R, n).astype(np.int32)
    return br, best["panels"].astype(np.int32)


def as_csr(x, device="cuda", long_row_threshold=256) -> CSRMatrix:
    if isinstance(x, (CSRMatrix, RowBlockedCSR)) or hasattr(x, "dist_spmm"):      # device CSR (row-blocked / row-partitioned too)
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


class RowBlockedCSR:
    """A sparse matrix too large for int32 CSR offsets (>= 2^31 non-zeros), kept as consecutive row blocks that
    each fit: the smoothed features X_conv = A_hat * X of a Twitter-World-sized input (main.py:528-530).  Offers
    what the minibatch MLP needs from its training matrix: ``shape``, ``gather_rows_device`` (mlp.py:81-91)."""

    def __init__(self, blocks, shape):
        self.blocks = list(blocks)
        self.shape = (int(shape[0]), int(shape[1]))
        self.row_off = np.zeros(len(self.blocks) + 1, np.int64)
        np.cumsum([b.shape[0] for b in self.blocks], out=self.row_off[1:])
        assert self.row_off[-1] == self.shape[0]
        self.device = self.blocks[0].device
        self.long_row_threshold = self.blocks[0].long_row_threshold

    @property
    def nnz(self):
        return int(sum(b.nnz for b in self.blocks))

    def astype(self, dtype):
        if np.dtype(dtype) != np.float32:
            raise ValueError("the B200 hot path is float32 only")
        return self

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.vstack([b.to_scipy() for b in self.blocks]).tocsr()

    def gather_rows_device(self, rows):
        """self[rows, :] as one CSRMatrix (rows in the requested order; duplicates allowed)"""
        rows = np.ascontiguousarray(np.asarray(rows), dtype=np.int64)
        if len(rows) and (rows.min() < 0 or rows.max() >= self.shape[0]):
            raise IndexError("row index out of range for a matrix with %d rows" % self.shape[0])
        owner = np.searchsorted(self.row_off, rows, side="right") - 1
        order = np.argsort(owner, kind="stable")
        parts, n_prev = [], 0
        for b, blk in enumerate(self.blocks):
            sel = order[owner[order] == b]
            if len(sel) == 0:
                continue
            parts.append(blk.gather_rows_device((rows[sel] - self.row_off[b]).astype(np.int32)))
        if len(parts) == 1 and np.array_equal(order, np.arange(len(rows))):
            return parts[0]
        ips, off = [], 0
        for p_ in parts:
            ips.append(p_.indptr[:-1].to(torch.int64) + off)
            off += p_.nnz
        if off >= 2 ** 31 - 1:
            raise ValueError("gathered rows do not fit int32 CSR offsets")
        ip = torch.cat(ips + [torch.tensor([off], dtype=torch.int64, device=self.device)]).to(torch.int32)
        cat = CSRMatrix(ip, torch.cat([p_.indices for p_ in parts]), torch.cat([p_.data for p_ in parts]),
                        (len(rows), self.shape[1]), long_row_threshold=self.long_row_threshold)
        inv = np.empty(len(rows), np.int32)
        inv[order] = np.arange(len(rows), dtype=np.int32)        # block-grouped position of every requested row
        return cat.gather_rows_device(inv)


def spgemm(A: "CSRMatrix", B: "CSRMatrix", a_values=None, max_block_nnz=2 ** 31 - 2):
    """C = A * B for two device CSR matrices, every entry summed in scipy's csr_matmat order, rows emitted
    with ascending columns (gcg_spgemm_count_csr / gcg_spgemm_fill_csr_f32).  ``a_values``: optional float64 device tensor replacing A.data -- the reference
    multiplies a FLOAT64 A_hat into X (main.py:522-530): float64 sums, one float32 rounding.  With float32
    values the sums are float32, scipy's rule for float32 * float32.
    A product with more than ``max_block_nnz`` non-zeros (int32 CSR offsets; Twitter-World smoothing) is returned
    as a ``RowBlockedCSR`` of consecutive row blocks, each filled by its own launch; otherwise a ``CSRMatrix``."""
    if A.shape[1] != B.shape[0]:
        raise ValueError("spgemm: A is %s but B is %s" % (A.shape, B.shape))
    dev = A.device
    L = _lib.lib()
    n, V = A.shape[0], B.shape[1]
    av = A.data if a_values is None else a_values
    if av.dtype not in (torch.float32, torch.float64) or av.numel() != A.nnz or not av.is_contiguous():
        raise TypeError("spgemm: a_values must be a contiguous float32/float64 tensor with A.nnz elements")
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    wsb = int(L.gcg_spgemm_workspace_bytes(V))
    ws = torch.empty(wsb + 256, dtype=torch.uint8, device=dev)
    wp = ws.data_ptr() + (-ws.data_ptr()) % 256
    row_nnz = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    _lib.check(L.gcg_spgemm_count_csr(n, V, A.indptr.data_ptr(), A.indices.data_ptr(), B.indptr.data_ptr(),
                                      B.indices.data_ptr(), 0, row_nnz.data_ptr(), wp, wsb, stream),
               "gcg_spgemm_count_csr")
    cum = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(row_nnz[:n], 0, out=cum[1:])
    total = int(cum[-1].item())
    # consecutive row blocks whose output fits the offset type (one block in the common case)
    bounds = [0]
    if total > max_block_nnz:
        cum_h = cum.cpu().numpy()
        while bounds[-1] < n:
            r0 = bounds[-1]
            r1 = int(np.searchsorted(cum_h, cum_h[r0] + max_block_nnz, side="right")) - 1
            if r1 <= r0:
                raise _lib.GcgError("spgemm: row %d alone has more than %d non-zeros" % (r0, max_block_nnz))
            bounds.append(min(r1, n))
    else:
        bounds.append(n)
    blocks = []
    esz = 8 if av.dtype == torch.float64 else 4
    for r0, r1 in zip(bounds[:-1], bounds[1:]):
        c_ip = (cum[r0:r1 + 1] - cum[r0]).contiguous()
        nnz = int(c_ip[-1].item())
        c_ix = torch.empty(nnz, dtype=torch.int32, device=dev)
        c_d = torch.empty(nnz, dtype=torch.float32, device=dev)
        # rows [r0, r1) of A: the same CSR arrays entered at row r0 (indptr holds absolute offsets)
        _lib.check(L.gcg_spgemm_fill_csr_f32(r1 - r0, V, A.indptr.data_ptr() + 4 * r0, A.indices.data_ptr(), av.data_ptr(),
                                             int(av.dtype == torch.float64), B.indptr.data_ptr(), B.indices.data_ptr(),
                                             B.data.data_ptr(), c_ip.data_ptr(), c_ix.data_ptr(), c_d.data_ptr(), wp, wsb,
                                             stream), "gcg_spgemm_fill_csr_f32")
        blocks.append(CSRMatrix(c_ip.to(torch.int32), c_ix, c_d, (r1 - r0, V)))
    del esz
    if len(blocks) == 1:
        return blocks[0]
    return RowBlockedCSR(blocks, (n, V))


def spgemm_pattern(A: "CSRMatrix", B: "CSRMatrix", drop_diagonal=False) -> "CSRMatrix":
    """Sorted column pattern of A * B (values ignored; the result's data are ones), optionally without the
    diagonal -- gcg_spgemm_count_csr / gcg_spgemm_fill_pattern_csr (graph projection, data.py:226-250)."""
    if A.shape[1] != B.shape[0]:
        raise ValueError("spgemm_pattern: A is %s but B is %s" % (A.shape, B.shape))
    dev = A.device
    L = _lib.lib()
    n, V = A.shape[0], B.shape[1]
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    wsb = int(L.gcg_spgemm_workspace_bytes(V))
    ws = torch.empty(wsb + 256, dtype=torch.uint8, device=dev)
    wp = ws.data_ptr() + (-ws.data_ptr()) % 256
    row_nnz = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    dd = int(bool(drop_diagonal))
    _lib.check(L.gcg_spgemm_count_csr(n, V, A.indptr.data_ptr(), A.indices.data_ptr(), B.indptr.data_ptr(),
                                      B.indices.data_ptr(), dd, row_nnz.data_ptr(), wp, wsb, stream),
               "gcg_spgemm_count_csr")
    c_ip = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(row_nnz[:n], 0, out=c_ip[1:])
    nnz = int(c_ip[-1].item())
    if nnz >= 2 ** 31 - 1:
        raise _lib.GcgError("spgemm_pattern: the product has %d non-zeros (int32 CSR offsets)" % nnz)
    c_ix = torch.empty(nnz, dtype=torch.int32, device=dev)
    _lib.check(L.gcg_spgemm_fill_pattern_csr(n, V, A.indptr.data_ptr(), A.indices.data_ptr(), B.indptr.data_ptr(),
                                             B.indices.data_ptr(), dd, c_ip.data_ptr(), c_ix.data_ptr(), wp, wsb,
                                             stream), "gcg_spgemm_fill_pattern_csr")
    return CSRMatrix(c_ip.to(torch.int32), c_ix, torch.ones(nnz, dtype=torch.float32, device=dev), (n, V))


def smooth_features(H, X, device="cuda", max_block_nnz=2 ** 31 - 2):
    """`X_conv = H * X; X_conv = X_conv.tocsr().astype('float32')` (main.py:528-530, tensormain.py:112-114).
    ``H``: scipy sparse A_hat (float64 as the reference builds it, or float32); ``X``: scipy CSR / CSRMatrix."""
    import scipy.sparse as sp
    Hs = sp.csr_matrix(H)
    A = CSRMatrix.from_scipy(Hs, device=device, sort_indices=False)
    a_vals = None
    if Hs.dtype == np.float64:
        a_vals = torch.from_numpy(np.ascontiguousarray(Hs.data)).to(A.device)
    B = X if isinstance(X, CSRMatrix) else CSRMatrix.from_scipy(X, device=device, sort_indices=False)
    return spgemm(A, B, a_values=a_vals, max_block_nnz=max_block_nnz)


def _min_block_cols():
    return int(_os.environ.get("GCG_XT_BLOCK_MIN_COLS", "1024"))     # tests lower it to block small inputs


class HeadSplit:
    """X = [dense head | sparse tail] by vocabulary term, for X.W1 and X^T.dZ1 at Twitter scale.

    Term frequencies are Zipf distributed: the ``k_head`` most frequent of 500 k terms hold ~36 % of X's
    non-zeros (k_head = 256, measured on the Twitter-World-shaped workload), i.e. those columns form a
    [N, 256] block that is ~15 % dense.  As a gather every one of its non-zeros pulls a 2.4 KB row of W1
    through L2; as a dense block it is a tensor-core GEMM (tcgen05 3xTF32) that reads X_head once:
        X.W1      = X_head . W1[top, :]  +  X_tail . W1              (GEMM + SpMM, accumulate)
        X^T.dZ1   = rows `top`: X_head^T . dZ1 (GEMM, split-K);  all other rows: X_tail^T . dZ1 (SpMM)
    The split is built once per fit (X is constant, mlpconv.py:169,294); index bookkeeping (frequency count,
    masked selection of the tail entries) uses torch on the device -- one-off data preparation, not the hot
    path.  Sums are re-associated (head first, then tail in CSR order): float32 rounding-level differences
    from the single-pass product, inside the north_star tolerance (tests/test_gpu_layers.py)."""

    @staticmethod
    def top_terms(indices, V, k_head):
        """ids of the ``k_head`` most frequent columns (ascending id; ties -> lower id), as an int64 device tensor"""
        df = torch.bincount(indices.to(torch.int64), minlength=V)
        top = torch.argsort(df, descending=True, stable=True)[:int(min(k_head, V))]
        return torch.sort(top).values

    def __init__(self, X: "CSRMatrix", k_head=256, top=None):
        """``top``: the head's term ids when they are decided elsewhere -- in multi-GPU runs every rank must use the
        SAME terms (chosen from the global X), because the head rows of dW1 are summed over ranks as one block."""
        from . import ops
        N, V = X.shape
        dev = X.device
        cols = X.indices.to(torch.int64)
        if top is None:
            top = self.top_terms(X.indices, V, k_head)     # ascending term id: the head keeps the column order
        top = top.to(device=dev, dtype=torch.int64)
        k_head = int(top.numel())
        pos = torch.full((V,), -1, dtype=torch.int64, device=dev)
        pos[top] = torch.arange(k_head, device=dev)
        deg = (X.indptr[1:] - X.indptr[:-1]).to(torch.int64)
        rows = torch.repeat_interleave(torch.arange(N, device=dev, dtype=torch.int64), deg)
        hp = pos[cols]
        in_head = hp >= 0
        self.k_head = k_head
        self.top = top
        self.top_i32 = top.to(torch.int32)
        self.head_nnz = int(in_head.sum().item())
        self.head_fraction = self.head_nnz / max(1, X.nnz)
        Xh = ops.alloc_mat(N, k_head, dev, zero=True)
        Xh[rows[in_head], hp[in_head]] = X.data[in_head]
        self.Xh = Xh
        keep = ~in_head
        t_deg = torch.zeros(N, dtype=torch.int64, device=dev).index_add_(0, rows[keep], torch.ones_like(rows[keep]))
        t_ip = torch.zeros(N + 1, dtype=torch.int64, device=dev)
        torch.cumsum(t_deg, 0, out=t_ip[1:])
        self.tail = CSRMatrix(t_ip.to(torch.int32), X.indices[keep].contiguous(), X.data[keep].contiguous(), (N, V),
                              long_row_threshold=X.long_row_threshold)
        del rows, hp, in_head, keep, cols
        self._split = None        # tf32 hi / lo copies of X_head (constant): made on first use
        self._w_head = None
        self._g_head = None

    def head_split(self):
        from . import ops
        if self._split is None and ops.gemm_uses_tensor_cores(self.Xh.shape[0], 600, self.k_head):
            self._split = ops.tf32_split(self.Xh)
        return self._split

    def product(self, W, out, bias=None, act="identity"):
        """out = act(X . W + bias)"""
        from . import ops
        F = W.shape[1]
        if self._w_head is None or self._w_head.shape != (self.k_head, F):
            self._w_head = ops.alloc_mat(self.k_head, F, W.device)
        ops.gather_rows(W, self.top_i32, out=self._w_head)
        if bias is None and act in ("identity", "linear", None):
            # no epilogue (the graph-conv layer applies bias / activation after A_hat): tail first through the lean
            # SpMM, head added by the GEMM's beta = 1 epilogue
            ops.spmm(self.tail, W, out=out)
            return ops.gemm(self.Xh, self._w_head, out=out, beta=1.0, a_split=self.head_split())
        ops.gemm(self.Xh, self._w_head, out=out, a_split=self.head_split())
        return ops.spmm(self.tail, W, out=out, accumulate=True, bias=bias, act=act)

    def transpose_product_head(self, dZ, dW, reduce=None):
        """dW[top, :] = X_head^T . dZ (the other rows of dW come from the tail's transpose product)"""
        from . import ops
        F = dZ.shape[1]
        if self._g_head is None or self._g_head.shape != (self.k_head, F):
            self._g_head = ops.alloc_mat(self.k_head, F, dZ.device)
        ops.gemm(self.Xh, dZ, transA=True, out=self._g_head, a_split=self.head_split())
        if reduce is not None:
            reduce(self._g_head).wait()
        return ops.put_rows(self._g_head, self.top_i32, dW)


class BlockedRows:
    """X^T.dZ with the frequent rows processed one column (document) block at a time.

    ``XT`` is the CSR of X^T [V, N].  Rows with at least ``heavy_factor * n_blocks`` non-zeros ("heavy"
    vocabulary terms; with Zipf-distributed terms they hold ~90 % of the non-zeros) are cut into
    ``n_blocks`` column blocks of ``block_cols`` documents (gcg_csr_split_colblocks_host): for one block all
    heavy rows gather from the same block_cols x F x 4 bytes of dZ, which stay L2-resident, and accumulate
    into a compact [n_heavy, F] buffer (empty (row, block) pairs are skipped by the kernel).  The light
    rows are done in one ordinary pass.  Result identical up to summation order inside the heavy rows."""

    @staticmethod
    def heavy_rows(row_nnz, n_cols, F, block_mb, heavy_factor):
        """the rule that makes a row of X^T 'heavy': at least ``heavy_factor`` non-zeros per document block"""
        block_cols = max(_min_block_cols(), int(block_mb * (1 << 20) // (4 * max(F, 1))))
        n_blocks = int(-(-n_cols // block_cols))
        return np.flatnonzero(np.asarray(row_nnz) >= heavy_factor * n_blocks).astype(np.int32)

    def __init__(self, XT: CSRMatrix, F, block_mb=32, heavy_factor=4, heavy_ids=None):
        """``heavy_ids``: the heavy rows when they are decided elsewhere -- in multi-GPU runs every rank must use the
        SAME set (from the global term frequencies): the compact heavy-row buffer is summed over ranks as one block."""
        ip, ix, d = XT._host_arrays()
        V, N = XT.shape
        self.shape = XT.shape
        self.device = XT.device
        block_cols = max(_min_block_cols(), int(block_mb * (1 << 20) // (4 * max(F, 1))))
        self.n_blocks = int(-(-N // block_cols))
        self.block_cols = block_cols
        lens = np.diff(ip)
        if heavy_ids is None:
            heavy = lens >= heavy_factor * self.n_blocks
        else:
            heavy = np.zeros(V, bool)
            heavy[np.asarray(heavy_ids, dtype=np.int64)] = True
        self.heavy_ids = np.flatnonzero(heavy).astype(np.int32)
        n_sel = len(self.heavy_ids)
        self.n_heavy = n_sel
        self.heavy_nnz_fraction = float(lens[heavy].sum()) / max(1, int(lens.sum()))
        if heavy_ids is None and (n_sel == 0 or self.n_blocks <= 1):
            self.light = XT
            self.blocks = []
            return
        # light rows: heavy rows emptied
        l_lens = np.where(heavy, 0, lens)
        l_ip = np.zeros(V + 1, np.int32)
        np.cumsum(l_lens, out=l_ip[1:])
        keep = np.repeat(~heavy, lens)
        self.light = CSRMatrix.from_host((l_ip, np.ascontiguousarray(ix[keep]), np.ascontiguousarray(d[keep])),
                                         (V, N), self.device, XT.long_row_threshold)
        del keep
        L = _lib.lib()
        nb = self.n_blocks
        o_ip = np.empty(nb * (n_sel + 1), np.int32)
        o_off = np.empty(nb + 1, np.int64)
        _lib.check(L.gcg_csr_split_colblocks_host(_np_ptr(ip), _np_ptr(ix), _np_ptr(d), _np_ptr(self.heavy_ids), n_sel,
                                                  block_cols, nb, _np_ptr(o_ip), _np_ptr(o_off), None, None),
                   "gcg_csr_split_colblocks_host")
        tot = int(o_off[-1])
        o_ix = np.empty(tot, np.int32)
        o_d = np.empty(tot, np.float32)
        _lib.check(L.gcg_csr_split_colblocks_host(_np_ptr(ip), _np_ptr(ix), _np_ptr(d), _np_ptr(self.heavy_ids), n_sel,
                                                  block_cols, nb, _np_ptr(o_ip), _np_ptr(o_off), _np_ptr(o_ix), _np_ptr(o_d)),
                   "gcg_csr_split_colblocks_host")
        dev_ix = torch.from_numpy(o_ix).to(self.device)
        dev_d = torch.from_numpy(o_d).to(self.device)
        dev_ip = torch.from_numpy(o_ip).to(self.device)
        self.blocks = []
        for b in range(nb):
            a, e = int(o_off[b]), int(o_off[b + 1])
            if e == a:
                continue
            hip = o_ip[b * (n_sel + 1):(b + 1) * (n_sel + 1)]
            blk = CSRMatrix(dev_ip[b * (n_sel + 1):(b + 1) * (n_sel + 1)], dev_ix[a:e], dev_d[a:e], (n_sel, N),
                            host=(hip, None, None), long_row_threshold=XT.long_row_threshold)
            # the block of dZ rows it gathers from is L2-resident by construction: register-gather kernel (L1-cached
            # loads) unless GCG_XT_BLOCK_KERNEL=stream asks for the streaming one (accumulate-only epilogue)
            blk.spmm_mode = _os.environ.get("GCG_XT_BLOCK_KERNEL", "gather")
            self.blocks.append(blk)
        self.heavy_dev = torch.from_numpy(self.heavy_ids.astype(np.int64)).to(self.device)
        self.heavy_i32 = self.heavy_dev.to(torch.int32)
        self._tmp = None
        # ONE matrix of all (document block, heavy row) pieces, block-major: row b*n_heavy + r holds the non-zeros of
        # heavy row r inside document block b.  A single balanced streaming SpMM over it (its spans run block by block,
        # so the block of dZ rows being gathered stays L2-resident) writes per-block partial rows; gcg_sum_slabs_f32
        # adds them in block order.  Replaces n_blocks launches that each re-read and re-wrote the heavy-row buffer.
        self.stacked = None
        if _os.environ.get("GCG_XT_STACKED", "1") != "0" and F % 4 == 0 and (n_sel * nb) < 2 ** 31 - 2:
            s_ip = np.empty(nb * n_sel + 1, np.int64)
            for b in range(nb):
                s_ip[b * n_sel:(b + 1) * n_sel] = o_ip[b * (n_sel + 1):(b + 1) * (n_sel + 1) - 1].astype(np.int64) + o_off[b]
            s_ip[-1] = tot
            s_ip32 = s_ip.astype(np.int32)
            self.stacked = CSRMatrix(torch.from_numpy(s_ip32).to(self.device), dev_ix, dev_d, (nb * n_sel, N),
                                     host=(s_ip32, None, None), long_row_threshold=XT.long_row_threshold)
            self.stacked.spmm_mode = _os.environ.get("GCG_XT_STACKED_KERNEL", "stream")
            self._partial = None

    def _light_chunks(self, n_chunks):
        """row slices of ``light`` (plans over contiguous row ranges of the same CSR arrays)"""
        key = int(n_chunks)
        if getattr(self, "_chunks", None) is None or self._chunks[0] != key:
            V = self.light.shape[0]
            ip = self.light._host_arrays()[0]
            bounds = np.linspace(0, V, key + 1).astype(np.int64)
            parts = []
            for a, e in zip(bounds[:-1], bounds[1:]):
                hip = np.ascontiguousarray(ip[a:e + 1])
                m = CSRMatrix(self.light.indptr[a:e + 1], self.light.indices, self.light.data, (int(e - a), self.light.shape[1]),
                              host=(hip, None, None), long_row_threshold=self.light.long_row_threshold)
                m.spmm_mode = getattr(self.light, "spmm_mode", None)
                parts.append((int(a), int(e), m))
            self._chunks = (key, parts)
        return self._chunks[1]

    def product(self, dZ, out, reduce=None, n_chunks=4):
        """out[V, F] = X^T . dZ.  ``reduce(tensor)`` (multi-GPU: an asynchronous all-reduce that returns a handle
        with ``wait()``) is applied to every finished piece as soon as it is computed -- the light rows in
        ``n_chunks`` row ranges, then the compact heavy-row buffer -- so that the summation over ranks of the
        largest gradient of the model overlaps the rest of this product instead of following it."""
        from . import ops
        pending = []
        if reduce is None:
            ops.spmm(self.light, dZ, out=out)
        else:
            for a, e, part in self._light_chunks(n_chunks):
                if e > a:
                    ops.spmm(part, dZ, out=out[a:e])
                    pending.append(reduce(out[a:e]))
        if self.blocks:
            F = dZ.shape[1]
            if self._tmp is None or self._tmp.shape != (self.n_heavy, F):
                self._tmp = ops.alloc_mat(self.n_heavy, F, dZ.device)
            if self.stacked is not None and self._tmp.is_contiguous():
                nb = self.stacked.shape[0] // self.n_heavy
                if self._partial is None or self._partial.shape != (nb * self.n_heavy, F):
                    self._partial = ops.alloc_mat(nb * self.n_heavy, F, dZ.device)
                ops.spmm(self.stacked, dZ, out=self._partial)
                ops.sum_slabs(self._partial, nb, self._tmp)
            else:
                for i, blk in enumerate(self.blocks):
                    ops.spmm(blk, dZ, out=self._tmp, accumulate=(i > 0))
            if reduce is not None:
                pending.append(reduce(self._tmp))
        for w in pending:
            w.wait()
        if self.blocks:
            ops.put_rows(self._tmp, self.heavy_i32, out)       # heavy rows are empty in `light`: plain placement
        return out
