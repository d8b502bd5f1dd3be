"""Thin Python wrappers over the libgcg.so C ABI (include/gcg.h).

torch is used only for device memory and the current CUDA stream.  Every
function raises if its inputs are not CUDA float32 tensors -- there is no CPU
or PyTorch fallback for the hot ops.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .sparse import CSRMatrix

# L2 budget used to pick the SpMM column-panel width: one panel of the gathered
# operand (n_cols * panel * 4 bytes) should stay resident in the 126 MB L2.
# Measured on B200 (profiles/r01_spmm_twitter_us.md): a 57.6 MB panel only reaches a 43% L2 hit
# rate (the two L2 partitions replicate lines read from both dies) and the re-read CSR arrays cost
# more than the hits save, so panel mode is OFF by default (budget 0) and selected explicitly.
L2_PANEL_BUDGET_BYTES = int(os.environ.get("GCG_L2_PANEL_BUDGET", 0))
_FORCE_PANEL = os.environ.get("GCG_SPMM_PANEL")          # experiments: force panel_cols
# which SpMM kernel family runs when the caller does not say: "auto" = the nnz-balanced streaming kernel
# (csrc/gcg_spmm_stream.cu) for matrices with >= STREAM_MIN_NNZ non-zeros, the register-gather kernel below that
# (measured on B200, Twitter-World A_hat.H at F = 600: 7.65 ms streaming vs 12.5 ms register-gather)
_SPMM_MODE = os.environ.get("GCG_SPMM_MODE", "auto")     # "auto" | "stream" | "gather"
STREAM_MIN_NNZ = int(os.environ.get("GCG_STREAM_MIN_NNZ", 1 << 20))
_GEMM_MODE = os.environ.get("GCG_GEMM_MODE", "auto")     # "fma" | "tf32x3" | "tf32" | "auto"
# OPT-IN (default 0 = off: every tensor-core product is 3xTF32, i.e. fp32-equivalent).  With a value K0 > 0, weight-
# gradient contractions (TN products, K = number of nodes / targets >= K0) run ONE tf32 pass on the round-to-nearest
# hi copies both operands already have: the input rounding (relative 4e-4 per product, unbiased) averages over
# K >= 65536 terms to below the fp32 accumulation noise of the 3-pass product itself.  Measured, reported separately
# (profiles/r02_gemm_notes.md); never on in the numbers bench.py reports as "dtype": "f32".
_LONGK_TF32 = int(os.environ.get("GCG_GEMM_LONGK_TF32", "0"))

_tc_available = None


def set_gemm_mode(mode):
    """"auto" (tcgen05 3xTF32 for contraction-bound shapes, FFMA otherwise) | "fma" | "tf32x3" | "tf32"."""
    global _GEMM_MODE
    if mode != "auto" and mode not in _lib.GEMM_MODE:
        raise ValueError("unknown gemm mode %r" % (mode,))
    _GEMM_MODE = mode


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# GCG_NVTX=1: every libgcg call below is wrapped in an NVTX range named after the reference operation it replaces
# (visible in nsys / ncu --nvtx); off by default (two extra host calls per launch)
_NVTX = os.environ.get("GCG_NVTX") == "1"


def _nvtx(name):
    def deco(fn):
        if not _NVTX:
            return fn

        def inner(*a, **k):
            torch.cuda.nvtx.range_push(name)
            try:
                return fn(*a, **k)
            finally:
                torch.cuda.nvtx.range_pop()
        inner.__name__, inner.__doc__ = fn.__name__, fn.__doc__
        return inner
    return deco


def round_up(x, m):
    return (x + m - 1) // m * m


def alloc_mat(n_rows, n_cols, device, zero=False):
    """[n_rows, n_cols] float32 view whose leading dimension is a multiple of 4
    floats (16-byte aligned rows) so that every kernel takes its 128-bit path."""
    ld = round_up(max(int(n_cols), 1), 4)
    buf = (torch.zeros if zero else torch.empty)((int(n_rows), ld), dtype=torch.float32, device=device)
    return buf[:, :n_cols] if ld != n_cols else buf


def _mat(t, name="tensor"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("%s must be a CUDA tensor (no CPU fallback)" % name)
    if t.dtype != torch.float32 or t.dim() != 2:
        raise TypeError("%s must be a 2-D float32 tensor" % name)
    if t.shape[1] > 1 and t.stride(1) != 1:
        raise ValueError("%s must be row-major (unit column stride)" % name)
    ld = t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])
    return C.c_void_p(t.data_ptr()), int(ld)


def _vec(t, name="tensor", dtype=torch.float32):
    if t is None:
        return None
    if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise TypeError("%s must be a contiguous CUDA %s tensor" % (name, dtype))
    return C.c_void_p(t.data_ptr())


# ------------------------------------------------------------------ scratch
class _Scratch:
    """One growing scratch buffer per (device, stream): all libgcg calls on a
    stream are ordered, so consecutive ops can share it."""

    def __init__(self, keep_outgrown=False):
        self.bufs = {}
        # an epoch program holds the pointers it was recorded with: buffers a later call outgrew stay allocated
        self.keep_outgrown = keep_outgrown
        self.outgrown = []

    def get(self, nbytes, device):
        if nbytes <= 0:
            return None, 0
        key = (device.index, torch.cuda.current_stream(device).cuda_stream)
        buf = self.bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            if buf is not None and self.keep_outgrown:
                self.outgrown.append(buf)
            buf = torch.empty(round_up(int(nbytes * 1.25), 512), dtype=torch.uint8, device=device)
            self.bufs[key] = buf
        return C.c_void_p(buf.data_ptr()), buf.numel()

    def reserve(self, nbytes, device):
        self.get(nbytes, device)


scratch = _Scratch()


# -------------------------------------------------------------------- epoch
class EpochProgram:
    """gcg_epoch (include/gcg.h): the libgcg calls of one f_train recorded once and replayed from C++ -- the
    counterpart of the compiled ``f_train = theano.function(...)`` of mlpconv.py:265 that mlpconv.py:295 calls
    once per epoch.  ``with prog.record(): step()`` executes the step AND records it; ``prog.run()`` enqueues the
    whole epoch on the current stream with one C call (CUDA-graph capturable).  The program owns the scratch
    its calls use; every other buffer belongs to the model, which must keep it in place (as for a CUDA graph)."""

    def __init__(self):
        L = _lib.lib()
        h = C.c_void_p()
        _lib.check(L.gcg_epoch_create(C.byref(h)), "gcg_epoch_create")
        self._h = h
        self._scratch = _Scratch(keep_outgrown=True)

    def record(self):
        import contextlib

        @contextlib.contextmanager
        def ctx():
            global scratch
            L = _lib.lib()
            _lib.check(L.gcg_epoch_record_begin(self._h), "gcg_epoch_record_begin")
            outer, scratch = scratch, self._scratch
            try:
                yield self
            finally:
                scratch = outer
                _lib.check(L.gcg_epoch_record_end(self._h), "gcg_epoch_record_end")
        return ctx()

    def run(self):
        _lib.check(_lib.lib().gcg_epoch_run(self._h, _stream()), "gcg_epoch_run")

    def __len__(self):
        return int(_lib.lib().gcg_epoch_size(self._h))

    def call_names(self):
        L = _lib.lib()
        return [L.gcg_epoch_call_name(self._h, i).decode() for i in range(len(self))]

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().gcg_epoch_destroy(self._h)
                self._h = None
        except Exception:
            pass


# --------------------------------------------------------------------- SpMM
STREAM_MIN_F = int(os.environ.get("GCG_STREAM_MIN_F", 272))


def auto_panel_cols(n_cols, F, nnz=0, prefer=None):
    """Kernel family for an SpMM the caller did not pin.  Measured on B200 (profiles/r02_spmm_stream.md), rows of
    the Twitter-World A_hat: the streaming kernel wins for wide operands gathered from HBM (F = 600: 7.7 vs 12.5
    ms, F = 1024: 14.0 vs 22.1, F = 300: 5.1 vs 10.2), ties at F = 256 / 152 and loses for narrow ones (F = 76:
    3.4 vs 2.7) and for operands
    that sit in L2 / L1 (the document blocks of X^T.dZ1: 1.9 vs 1.0 ms), which ``prefer="gather"`` marks."""
    if _FORCE_PANEL is not None:
        return int(_FORCE_PANEL)
    mode = _SPMM_MODE if _SPMM_MODE != "auto" else (prefer or "auto")
    if F <= 1024 and (mode == "stream" or (mode == "auto" and nnz >= STREAM_MIN_NNZ and F >= STREAM_MIN_F)):
        return -2
    if L2_PANEL_BUDGET_BYTES <= 0 or n_cols * F * 4 <= L2_PANEL_BUDGET_BYTES:
        return 0
    p = 16
    while p * 2 < F and n_cols * (p * 2) * 4 <= L2_PANEL_BUDGET_BYTES:
        p *= 2
    return p


@_nvtx("S.dot -> gcg_spmm_csr_f32")
def spmm(A: CSRMatrix, B, out=None, bias=None, act="identity", accumulate=False, gate=None, carry=None,
         conv_out=None, panel_cols=None):
    """out = epilogue(A @ B) -- gcg_spmm_csr_f32 (S.dot, lasagne_layers.py:26,65,67,84)."""
    if hasattr(A, "dist_spmm"):        # row-partitioned matrix: all-gather of B overlapped with the local block
        assert not accumulate
        return A.dist_spmm(B, out=out, bias=bias, act=act, gate=gate, carry=carry, conv_out=conv_out)
    L = _lib.lib()
    bp, ldb = _mat(B, "B")
    n, k = A.shape
    F = B.shape[1]
    if B.shape[0] != k:
        raise ValueError("spmm: A is %s but B has %d rows" % (A.shape, B.shape[0]))
    if out is None:
        out = alloc_mat(n, F, B.device)
    if out.shape != (n, F):
        raise ValueError("spmm: out has shape %s, expected %s" % (tuple(out.shape), (n, F)))
    if n == 0:
        return out
    cp, ldc = _mat(out, "out")
    gp = hp = vp = None
    ldg = ldh = ldv = 0
    if gate is not None:
        gp, ldg = _mat(gate, "gate")
        hp, ldh = _mat(carry, "carry")
        if conv_out is not None:
            vp, ldv = _mat(conv_out, "conv_out")
    if panel_cols is None:
        panel_cols = auto_panel_cols(k, F, A.nnz, getattr(A, "spmm_mode", None))
    ws, wsb = scratch.get(A.workspace_bytes(ldc), B.device)
    _lib.check(L.gcg_spmm_csr_f32(A.plan, bp, ldb, F, cp, ldc, _vec(bias, "bias"), _lib.act_code(act),
                                  int(bool(accumulate)), gp, ldg, hp, ldh, vp, ldv, int(panel_cols), ws, wsb,
                                  _stream()), "gcg_spmm_csr_f32")
    return out


# --------------------------------------------------------------------- GEMM
def gemm_mode_for(M, N, K):
    """Engine choice for a dense projection: tcgen05 3xTF32 when the contraction is
    large enough to be compute-bound on FFMA, else fused FMA tiles."""
    global _tc_available
    if _GEMM_MODE != "auto":
        return _lib.GEMM_MODE[_GEMM_MODE]
    if _tc_available is None:
        _tc_available = hasattr(_lib.lib(), "gcg_gemm_tc_available") and bool(_lib.lib().gcg_gemm_tc_available())
    if _tc_available and 2.0 * M * N * K >= 2e9:
        return _lib.GEMM_MODE["tf32x3"]
    return _lib.GEMM_MODE["fma"]


class SplitOperand:
    """tf32 hi/lo copies of a matrix (gcg_tf32_split_f32), reusable by every GEMM that reads it."""

    def __init__(self, src, hi=None, lo=None):
        sp, ld = _mat(src, "src")
        n, c = src.shape
        self.hi = alloc_mat(n, c, src.device) if hi is None else hi
        self.lo = alloc_mat(n, c, src.device) if lo is None else lo
        hp, ldh = _mat(self.hi, "hi")
        lp, ldl = _mat(self.lo, "lo")
        self.valid = (ldh == ld and ldl == ld and ld % 4 == 0)
        self.ptr, self.ld, self.shape = src.data_ptr(), ld, tuple(src.shape)
        if self.valid:
            _lib.check(_lib.lib().gcg_tf32_split_f32(sp, ld, n, hp, lp, _stream()), "gcg_tf32_split_f32")

    def matches(self, t, ld):
        return self.valid and t.data_ptr() == self.ptr and tuple(t.shape) == self.shape and ld == self.ld


def tf32_split(src, like=None):
    """Split ``src`` once for several tcgen05 GEMMs; ``like`` reuses a previous SplitOperand's buffers."""
    if like is not None and like.hi.shape == src.shape:
        return SplitOperand(src, like.hi, like.lo)
    return SplitOperand(src)


def gemm_uses_tensor_cores(M, N, K):
    return gemm_mode_for(M, N, K) in (_lib.GEMM_MODE["tf32x3"], _lib.GEMM_MODE["tf32x3_chained"])


@_nvtx("T.dot -> gcg_gemm_f32")
def gemm(A, B, out=None, transA=False, transB=False, beta=0.0, bias=None, act="identity", mask=None,
         mask_act="identity", mode=None, split_k=0, a_split=None, b_split=None, chained=False):
    """out = act(op(A) @ op(B) + beta*out + bias) [* act'(mask)] -- gcg_gemm_f32 (T.dot, lasagne_layers.py:82).
    ``chained``: when the tensor-core engine runs this product, keep its accumulation chains short
    (GCG_GEMM_TF32X3_CHAINED) -- for outputs whose terms cancel, i.e. the logits."""
    L = _lib.lib()
    ap, lda = _mat(A, "A")
    bp, ldb = _mat(B, "B")
    M, K = (A.shape[1], A.shape[0]) if transA else A.shape
    K2, N = (B.shape[1], B.shape[0]) if transB else B.shape
    if K != K2:
        raise ValueError("gemm: inner dimensions differ (%d vs %d)" % (K, K2))
    if out is None:
        if beta != 0.0:
            raise ValueError("gemm: beta != 0 needs an existing out")
        out = alloc_mat(M, N, A.device)
    if out.shape != (M, N):
        raise ValueError("gemm: out has shape %s, expected %s" % (tuple(out.shape), (M, N)))
    if M == 0 or N == 0:
        return out
    if K == 0:                    # empty contraction (e.g. a rank without target rows): op(A).op(B) = 0
        if beta == 0.0:
            out.zero_()
        elif beta != 1.0:
            out.mul_(beta)
        if bias is not None or act != "identity" or mask is not None:
            raise ValueError("gemm: K == 0 with an epilogue is not supported")
        return out
    cp, ldc = _mat(out, "out")
    mp, ldm = (None, 0) if mask is None else _mat(mask, "mask")
    auto_mode = mode is None
    if mode is None:
        mode = gemm_mode_for(M, N, K)
    elif isinstance(mode, str):
        mode = _lib.GEMM_MODE[mode]
    if chained and mode == _lib.GEMM_MODE["tf32x3"]:
        mode = _lib.GEMM_MODE["tf32x3_chained"]
    have_a = a_split is not None and a_split.matches(A, lda)
    have_b = b_split is not None and b_split.matches(B, ldb)
    if (auto_mode and mode == _lib.GEMM_MODE["tf32x3"] and transA and not transB and have_a and have_b
            and 0 < _LONGK_TF32 <= K):
        mode = _lib.GEMM_MODE["tf32"]                    # long weight-gradient contraction: see _LONGK_TF32
    x3 = mode in (_lib.GEMM_MODE["tf32x3"], _lib.GEMM_MODE["tf32x3_chained"], _lib.GEMM_MODE["tf32"])
    wbytes = L.gcg_gemm_workspace_bytes(int(transA), int(transB), M, N, K, mode, int(split_k))
    ws, wsb = scratch.get(wbytes, A.device)
    pre = [None, None, None, None]
    if x3:
        if have_a:
            pre[0], pre[1] = C.c_void_p(a_split.hi.data_ptr()), C.c_void_p(a_split.lo.data_ptr())
        if have_b:
            pre[2], pre[3] = C.c_void_p(b_split.hi.data_ptr()), C.c_void_p(b_split.lo.data_ptr())
    if pre[0] is not None or pre[2] is not None:
        _lib.check(L.gcg_gemm_presplit_f32(int(transA), int(transB), M, N, K, ap, lda, bp, ldb, cp, ldc, float(beta),
                                           _vec(bias, "bias"), _lib.act_code(act), mp, ldm, _lib.act_code(mask_act),
                                           mode, int(split_k), ws, wsb, _stream(), *pre), "gcg_gemm_presplit_f32")
        return out
    _lib.check(L.gcg_gemm_f32(int(transA), int(transB), M, N, K, ap, lda, bp, ldb, cp, ldc, float(beta),
                              _vec(bias, "bias"), _lib.act_code(act), mp, ldm, _lib.act_code(mask_act), mode,
                              int(split_k), ws, wsb, _stream()), "gcg_gemm_f32")
    return out


# ------------------------------------------------- epilogues / reductions
def colsum(X, out=None):
    L = _lib.lib()
    xp, ld = _mat(X, "X")
    n, F = X.shape
    if out is None:
        out = torch.empty(F, dtype=torch.float32, device=X.device)
    if n == 0 or F == 0:          # a rank that owns no rows: the sum over nothing (an empty tensor has a NULL pointer)
        return out.zero_()
    ws, wsb = scratch.get(L.gcg_colsum_workspace_bytes(n, F), X.device)
    _lib.check(L.gcg_colsum_f32(xp, ld, n, F, _vec(out, "out"), ws, wsb, _stream()), "gcg_colsum_f32")
    return out


def act_bwd(dA, A, act, out=None):
    L = _lib.lib()
    dp, ldd = _mat(dA, "dA")
    ap, lda = _mat(A, "A")
    if out is None:
        out = alloc_mat(dA.shape[0], dA.shape[1], dA.device)
    if dA.shape[0] == 0 or dA.shape[1] == 0:
        return out
    op, ldo = _mat(out, "out")
    _lib.check(L.gcg_act_bwd_f32(dp, ldd, ap, lda, op, ldo, dA.shape[0], dA.shape[1], _lib.act_code(act),
                                 _stream()), "gcg_act_bwd_f32")
    return out


def highway_mix(Hc, g, Hin, out=None):
    """out = g*Hc + (1-g)*Hin -- gcg_highway_fwd_f32 (normally fused into the SpMM epilogue)."""
    if out is None:
        out = alloc_mat(Hc.shape[0], Hc.shape[1], Hc.device)
    a = [_mat(t, nm) for t, nm in ((Hc, "Hc"), (g, "g"), (Hin, "Hin"), (out, "out"))]
    flat = [x for pair in a for x in pair]
    _lib.check(_lib.lib().gcg_highway_fwd_f32(*flat, Hc.shape[0], Hc.shape[1], _stream()), "gcg_highway_fwd_f32")
    return out


@_nvtx("highway gate backward -> gcg_highway_bwd_f32")
def highway_bwd(dO, g, Hc, Hin, act, dP=None, dGpre=None, dHin=None):
    L = _lib.lib()
    n, F = dO.shape
    dev = dO.device
    dP = alloc_mat(n, F, dev) if dP is None else dP
    dGpre = alloc_mat(n, F, dev) if dGpre is None else dGpre
    dHin = alloc_mat(n, F, dev) if dHin is None else dHin
    a = [_mat(t, nm) for t, nm in ((dO, "dO"), (g, "g"), (Hc, "Hc"), (Hin, "Hin"), (dP, "dP"),
                                   (dGpre, "dGpre"), (dHin, "dHin"))]
    flat = [x for pair in a for x in pair]
    _lib.check(L.gcg_highway_bwd_f32(*flat, n, F, _lib.act_code(act), _stream()), "gcg_highway_bwd_f32")
    return dP, dGpre, dHin


@_nvtx("softmax + categorical_crossentropy -> gcg_softmax_ce_f32")
def softmax_ce(logits, y=None, probs=None, grad=None, ce=None, hit=None, pred=None, denom=None):
    """Output head on gathered logits -- gcg_softmax_ce_f32 (mlpconv.py:216,223,227-233,252)."""
    L = _lib.lib()
    lp, ldl = _mat(logits, "logits")
    n, Cc = logits.shape
    pp, ldp = (None, 0) if probs is None else _mat(probs, "probs")
    gp, ldg = (None, 0) if grad is None else _mat(grad, "grad")
    if pred is not None and pred.dtype != torch.int64:
        raise TypeError("pred must be int64 (Theano argmax dtype)")
    _lib.check(L.gcg_softmax_ce_f32(lp, ldl, _vec(y, "y", torch.int32), n, Cc,
                                    float(denom if denom is not None else max(n, 1)), pp, ldp, gp, ldg,
                                    _vec(ce, "ce"), _vec(hit, "hit"), _vec(pred, "pred", torch.int64),
                                    _stream()), "gcg_softmax_ce_f32")


def sum_scaled(x, scale=1.0, out=None):
    if out is None:
        out = torch.empty(1, dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().gcg_sum_f32(_vec(x, "x"), x.numel(), float(scale), _vec(out, "out"), _stream()),
               "gcg_sum_f32")
    return out


def sum_slabs(src, n_slabs, out):
    """out = sum over the n_slabs equal slabs of ``src`` in slab order (gcg_sum_slabs_f32)."""
    slab = out.numel()
    if src.numel() != n_slabs * slab or not src.is_contiguous() or not out.is_contiguous():
        raise ValueError("sum_slabs: src must be %d contiguous slabs of out's size" % n_slabs)
    _lib.check(_lib.lib().gcg_sum_slabs_f32(C.c_void_p(src.data_ptr()), int(n_slabs), int(slab), C.c_void_p(out.data_ptr()),
                                            _stream()), "gcg_sum_slabs_f32")
    return out


def scatter_rows(G, pos_ptr, pos_idx, n_rows, out=None):
    L = _lib.lib()
    gp, ldg = _mat(G, "G")
    Cc = G.shape[1]
    if out is None:
        out = alloc_mat(n_rows, Cc, G.device)
    if n_rows == 0 or Cc == 0:
        return out
    if G.shape[0] == 0:           # no target rows on this rank: nothing is scattered, the gradient slab is zero
        return out.zero_()
    op, ldo = _mat(out, "out")
    _lib.check(L.gcg_scatter_rows_f32(gp, ldg, _vec(pos_ptr, "pos_ptr", torch.int32),
                                      _vec(pos_idx, "pos_idx", torch.int32), n_rows, Cc, op, ldo, _stream()),
               "gcg_scatter_rows_f32")
    return out


def gather_rows(X, idx, out=None):
    L = _lib.lib()
    xp, ldx = _mat(X, "X")
    n, Cc = idx.numel(), X.shape[1]
    if out is None:
        out = alloc_mat(n, Cc, X.device)
    op, ldo = _mat(out, "out")
    _lib.check(L.gcg_gather_rows_f32(xp, ldx, _vec(idx, "idx", torch.int32), n, Cc, op, ldo, _stream()),
               "gcg_gather_rows_f32")
    return out


def put_rows(src, idx, dst):
    """dst[idx[i], :] = src[i, :] for distinct ``idx`` (int32 device tensor) -- gcg_put_rows_f32."""
    sp, lds = _mat(src, "src")
    dp, ldd = _mat(dst, "dst")
    n, Cc = idx.numel(), src.shape[1]
    if src.shape[0] != n or dst.shape[1] != Cc:
        raise ValueError("put_rows: src is %s for %d indices, dst %s" % (tuple(src.shape), n, tuple(dst.shape)))
    _lib.check(_lib.lib().gcg_put_rows_f32(sp, lds, _vec(idx, "idx", torch.int32), n, Cc, dp, ldd, _stream()),
               "gcg_put_rows_f32")
    return dst


def pack_cols(src, P, Fp, dst):
    """[n, F] -> [P, n, Fp] column slices, zero padded (gcg_pack_cols_f32)."""
    sp, ld = _mat(src, "src")
    _lib.check(_lib.lib().gcg_pack_cols_f32(sp, ld, src.shape[0], src.shape[1], int(P), int(Fp),
                                            C.c_void_p(dst.data_ptr()), _stream()), "gcg_pack_cols_f32")
    return dst


def unpack_cols(src, P, Fp, dst):
    """[P, n, Fp] column slices -> [n, F] (gcg_unpack_cols_f32)."""
    dp, ld = _mat(dst, "dst")
    _lib.check(_lib.lib().gcg_unpack_cols_f32(C.c_void_p(src.data_ptr()), dst.shape[0], dst.shape[1], int(P), int(Fp),
                                              dp, ld, _stream()), "gcg_unpack_cols_f32")
    return dst


def scatter_positions(idx, n_rows, device):
    """CSR-like inverse of ``target_indices``: for node r the positions i with
    idx[i] == r, ascending (host, once per index vector)."""
    idx = np.asarray(idx, dtype=np.int64)
    order = np.argsort(idx, kind="stable").astype(np.int32)
    counts = np.bincount(idx, minlength=n_rows)
    ptr = np.zeros(n_rows + 1, np.int32)
    np.cumsum(counts, out=ptr[1:])
    return torch.from_numpy(ptr).to(device), torch.from_numpy(order).to(device)


# ---------------------------------------------------------------- optimiser
class Adam:
    """lasagne.updates.adam(lr=4e-3, 0.9, 0.999, 1e-8) (mlpconv.py:263) over a fixed
    parameter list, one fused launch (gcg_adam_step_f32).  ``reg`` holds the
    elastic-net coefficient per tensor (0 for biases; mlpconv.py:235-244)."""

    def __init__(self, params, grads, reg, lr=4e-3, beta1=0.9, beta2=0.999, eps=1e-8):
        self.params, self.grads = list(params), list(grads)
        n = len(self.params)
        for p, g in zip(self.params, self.grads):
            if not (p.is_contiguous() and g.is_contiguous() and p.numel() == g.numel()):
                raise ValueError("Adam needs contiguous parameters and gradients of equal size")
        dev = self.params[0].device
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.t = torch.zeros(2, dtype=torch.float32, device=dev)     # [t, a_t]
        self.reg_out = torch.zeros(1, dtype=torch.float32, device=dev)
        self.hyper = (float(lr), float(beta1), float(beta2), float(eps))
        arr = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])
        self._p, self._g, self._m, self._v = arr(self.params), arr(self.grads), arr(self.m), arr(self.v)
        self._sizes = (C.c_int64 * n)(*[p.numel() for p in self.params])
        self._reg = (C.c_float * n)(*[float(r) for r in reg])
        self._n = n
        self._wsbytes = int(_lib.lib().gcg_adam_workspace_bytes(n, self._sizes))
        self._ws = torch.empty(max(self._wsbytes, 4), dtype=torch.uint8, device=dev)

    def step(self):
        lr, b1, b2, eps = self.hyper
        _lib.check(_lib.lib().gcg_adam_step_f32(self._n, self._p, self._g, self._m, self._v, self._sizes,
                                                self._reg, lr, b1, b2, eps, C.c_void_p(self.t.data_ptr()),
                                                C.c_void_p(self.reg_out.data_ptr()),
                                                C.c_void_p(self._ws.data_ptr()), self._ws.numel(), _stream()),
                   "gcg_adam_step_f32")
        return self.reg_out


class ElasticNet:
    """0.5*c*(|W|_1 + |W|_2^2) summed over a fixed tensor list (mlpconv.py:235-245),
    for eval_loss where no update happens (gcg_elastic_net_f32)."""

    def __init__(self, params, reg):
        self.params = list(params)
        n = len(self.params)
        dev = self.params[0].device
        self._p = (C.c_void_p * n)(*[t.data_ptr() for t in self.params])
        self._sizes = (C.c_int64 * n)(*[p.numel() for p in self.params])
        self._reg = (C.c_float * n)(*[float(r) for r in reg])
        self._n = n
        wsb = int(_lib.lib().gcg_adam_workspace_bytes(n, self._sizes))
        self._ws = torch.empty(max(wsb, 4), dtype=torch.uint8, device=dev)
        self.out = torch.zeros(1, dtype=torch.float32, device=dev)

    def __call__(self):
        _lib.check(_lib.lib().gcg_elastic_net_f32(self._n, self._p, self._sizes, self._reg,
                                                  C.c_void_p(self.out.data_ptr()),
                                                  C.c_void_p(self._ws.data_ptr()), self._ws.numel(), _stream()),
                   "gcg_elastic_net_f32")
        return self.out


# ------------------------------------------------------------------- host
def kdtree_fit(points, bucket_size):
    """Region label per point (int64) and number of leaves -- gcg_kdtree_fit_host,
    bit-exact restatement of kdtree.py:84-118,126-147."""
    pts = np.ascontiguousarray(np.asarray(points, dtype=np.float64))
    if pts.ndim != 2:
        raise ValueError("points must be [n, dims]")
    n, dims = pts.shape
    labels = np.zeros(n, dtype=np.int64)
    nl = C.c_int64(0)
    _lib.check(_lib.lib().gcg_kdtree_fit_host(pts.ctypes.data_as(C.c_void_p), n, dims, int(bucket_size),
                                              labels.ctypes.data_as(C.c_void_p), C.byref(nl)),
               "gcg_kdtree_fit_host")
    return labels, int(nl.value)


def launch_count(reset=False):
    L = _lib.lib()
    n = int(L.gcg_launch_count())
    if reset:
        L.gcg_launch_count_reset()
    return n
