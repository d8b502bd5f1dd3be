"""The reference's layer API for the GCN propagation path, on B200.

Mirrors /root/reference/lasagne_layers.py:20-89 (byte-identical copies live in
mlpconv.py:27-95 and mlp.py:36-45): same class names, same constructor keywords
(``incoming, num_units, W=, b=, nonlinearity=, H=``) and the same
``get_output_for(input, **kwargs)`` entry point honouring ``target_indices``
(:81) and ``deterministic`` (:32,37).  The bodies call the hand-written sm_100a
kernels of libgcg.so instead of Theano's S.dot / T.dot; there is no Theano
graph, so ``get_output_for`` executes eagerly on CUDA tensors.

Because there is no symbolic autodiff either, every layer also implements
``backward(grad_output)`` (the products theano.grad derives at mlpconv.py:263),
storing parameter gradients in ``layer.grads``.

Extension (BASELINE.json north_star, not in the reference):
``HighwayConvolutionDenseLayer`` -- a conv layer whose output is gated with its
input, out = g*H' + (1-g)*H, g = sigmoid(H.W_g + b_g); the mix is fused into the
SpMM epilogue.  ``GraphConvLayer`` is the north_star's name for the conv layers.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .sparse import CSRMatrix, as_csr, is_sparse

# Forward projections H.W of the hidden layers on the tensor-core engine keep their accumulation chains short
# (GCG_GEMM_TF32X3_CHAINED, as the logits product does): after ~25 epochs the hidden activations of the Twitter
# shapes reach ~16 and the one-chain product sat at 1.14x the parity bound on 2 of 614,400 sampled values
# (profiles/r02_parity_notes.md); the gate product feeds a sigmoid and stays as it was (0.08x).  0 switches it off.
import os as _os
_FWD_CHAINED = _os.environ.get("GCG_FWD_CHAINED", "1") != "0"

# --------------------------------------------------------------------------- #
# lasagne.nonlinearities / lasagne.init look-alikes                            #
# --------------------------------------------------------------------------- #


class _NL(str):
    """A non-linearity name that can also be compared / passed like Lasagne's functions."""


class nonlinearities:
    rectify = _NL("rectify")
    tanh = _NL("tanh")
    sigmoid = _NL("sigmoid")
    softmax = _NL("softmax")
    identity = linear = _NL("identity")


def _nl_name(nl):
    if nl is None:
        return "identity"
    if isinstance(nl, str):
        name = str(nl)
    else:
        name = getattr(nl, "__name__", None)
    if name == "relu":
        name = "rectify"
    if name == "linear":
        name = "identity"
    if name not in ("rectify", "tanh", "sigmoid", "softmax", "identity"):
        raise ValueError("unsupported nonlinearity %r" % (nl,))
    return name


class GlorotUniform:
    """lasagne.init.GlorotUniform (mlpconv.py:208): U(-a, a), a = sqrt(6/(fan_in+fan_out))."""

    def __init__(self, gain=1.0):
        self.gain = gain

    def sample(self, shape, rng=None):
        rng = np.random if rng is None else rng
        a = self.gain * np.sqrt(6.0 / (shape[0] + shape[1]))
        return rng.uniform(-a, a, size=shape).astype(np.float32)


class Constant:
    def __init__(self, val=0.0):
        self.val = val

    def sample(self, shape, rng=None):
        return np.full(shape, self.val, dtype=np.float32)


class init:
    GlorotUniform = GlorotUniform
    Constant = Constant


def _to_param(spec, shape, device, rng=None):
    if spec is None:
        return None
    if hasattr(spec, "sample"):
        arr = spec.sample(shape, rng)
    elif isinstance(spec, torch.Tensor):
        t = spec.to(device=device, dtype=torch.float32).contiguous()
        assert tuple(t.shape) == tuple(shape), "parameter shape %s != %s" % (tuple(t.shape), shape)
        return t
    else:
        arr = np.asarray(spec, dtype=np.float32)
    assert tuple(arr.shape) == tuple(shape), "parameter shape %s != %s" % (arr.shape, shape)
    return torch.from_numpy(np.ascontiguousarray(arr)).to(device)


# --------------------------------------------------------------------------- #
# Layer protocol                                                               #
# --------------------------------------------------------------------------- #


class Layer:
    def __init__(self, incoming, name=None, device="cuda"):
        if isinstance(incoming, Layer):
            self.input_layer = incoming
            self.input_shape = incoming.output_shape
            self.device = incoming.device
        else:
            self.input_layer = None
            self.input_shape = tuple(incoming)
            self.device = torch.device(device)
        self.name = name
        self.params = []          # [(name, tensor, tags)]
        self.grads = {}
        self._buf = {}

    @property
    def output_shape(self):
        return self.get_output_shape_for(self.input_shape)

    def get_output_shape_for(self, input_shape):
        return input_shape

    def get_params(self, **tags):
        out = []
        for _, t, ptags in self.params:
            if all(ptags.get(k, False) == v for k, v in tags.items()):
                out.append(t)
        return out

    def get_output_for(self, input, **kwargs):
        raise NotImplementedError

    def backward(self, grad_output, **kwargs):
        raise NotImplementedError

    # persistent, shape-keyed buffers: nothing is allocated in the steady state,
    # which keeps the epoch CUDA-graph capturable
    def _mat(self, key, n_rows, n_cols, zero=False):
        b = self._buf.get(key)
        if b is None or b.shape != (n_rows, n_cols):
            b = ops.alloc_mat(n_rows, n_cols, self.device, zero=zero)
            self._buf[key] = b
        return b

    def _vecbuf(self, key, n, dtype=torch.float32):
        b = self._buf.get(key)
        if b is None or b.numel() != n or b.dtype != dtype:
            b = torch.empty(n, dtype=dtype, device=self.device)
            self._buf[key] = b
        return b


class InputLayer(Layer):
    """lasagne.layers.InputLayer (mlpconv.py:196-197)."""

    def __init__(self, shape, input_var=None, name=None, device="cuda"):
        super().__init__(shape, name=name, device=device)
        self.shape = tuple(shape)
        self.input_var = input_var

    def get_output_for(self, input, **kwargs):
        return input


class DenseLayer(Layer):
    """lasagne.layers.DenseLayer: W [num_inputs, num_units] (regularizable, trainable),
    b [num_units] (trainable); act(input.W + b)."""

    def __init__(self, incoming, num_units, W=None, b=Constant(0.0), nonlinearity=nonlinearities.rectify,
                 name=None, rng=None, **kwargs):
        super().__init__(incoming, name=name, **kwargs)
        self.num_units = int(num_units)
        self.nonlinearity = _nl_name(nonlinearity)
        num_inputs = int(np.prod(self.input_shape[1:]))
        self.num_inputs = num_inputs
        W = GlorotUniform() if W is None else W
        self.W = _to_param(W, (num_inputs, self.num_units), self.device, rng)
        self.b = _to_param(b, (self.num_units,), self.device, rng)
        self.params.append(("W", self.W, dict(trainable=True, regularizable=True)))
        if self.b is not None:
            self.params.append(("b", self.b, dict(trainable=True, regularizable=False)))

    def get_output_shape_for(self, input_shape):
        return (input_shape[0], self.num_units)

    def get_output_for(self, input, **kwargs):
        """act(input.W + b).  A softmax layer (the MLP head, mlp.py:183-185) returns probabilities, or
        the logits with ``logits=True`` (the training step feeds them to the fused softmax/CE head)."""
        n = input.shape[0]
        self._in = input
        fused_act = "identity" if self.nonlinearity == "softmax" else self.nonlinearity
        out = ops.gemm(input, self.W, bias=self.b, act=fused_act, out=self._mat(("out", n), n, self.num_units))
        self._out = out
        return self._softmax_or(out, kwargs)

    def _softmax_or(self, out, kwargs):
        if self.nonlinearity == "softmax" and not kwargs.get("logits", False):
            n = out.shape[0]
            probs = self._mat(("probs", n), n, self.num_units)
            ops.softmax_ce(out, probs=probs)
            return probs
        return out

    def backward(self, grad_output, preact=False, input_mask=None, need_input_grad=True, **kwargs):
        """T.dot's gradients: dW = in^T.dP, db = colsum(dP), d(in) = dP.W^T.  ``grad_output`` is the grad wrt
        the LOGITS for a softmax layer; ``input_mask=(A_prev, act)`` fuses the previous layer's act'."""
        n = grad_output.shape[0]
        if preact or self.nonlinearity in ("softmax", "identity"):
            dP = grad_output
        else:
            dP = ops.act_bwd(grad_output, self._out, self.nonlinearity, out=self._mat(("dP", n), n, self.num_units))
        if self.b is not None:
            ops.colsum(dP, out=self._grad("b", self.b))
        ops.gemm(self._in, dP, transA=True, out=self._grad("W", self.W))
        if not need_input_grad:
            return None
        mk, mact = (None, "identity") if input_mask is None else input_mask
        return ops.gemm(dP, self.W, transB=True, mask=mk, mask_act=mact,
                        out=self._mat(("dIn", n), n, self.num_inputs))

    def _split(self, key, t, other_dim):
        """tf32 hi/lo split of an activation that several tcgen05 GEMMs of this step read (None when the
        contraction is small enough for the FFMA engine)."""
        if not ops.gemm_uses_tensor_cores(t.shape[0], t.shape[1], other_dim):
            return None
        if not hasattr(self, "_splits"):
            self._splits = {}
        sp = ops.tf32_split(t, like=self._splits.get(key))
        self._splits[key] = sp
        return sp

    def _grad(self, key, like):
        g = self.grads.get(key)
        if g is None:
            g = torch.zeros_like(like)
            self.grads[key] = g
        return g


def _is_big(X, F):
    import os
    if os.environ.get("GCG_X_FORCE_BIG") == "1":      # tests: exercise the Twitter-scale code paths on small inputs
        return True
    return X.nnz >= (1 << 24) and X.shape[0] * F * 4 > (256 << 20)


def _head_split(layer, X, F):
    """sparse.HeadSplit of a Twitter-scale X (None for small inputs or GCG_X_HEAD=0): the most frequent terms as
    a dense block for the tensor cores, the rest as CSR.  Built once and kept on the layer."""
    import os
    from .sparse import HeadSplit
    k = int(os.environ.get("GCG_X_HEAD", "256"))
    plan = getattr(layer, "_x_plan", None)         # multi-GPU: decisions taken once from the GLOBAL X (dist.py)
    if plan is not None:
        if plan.get("top") is None:
            return None
    elif k <= 0 or not _is_big(X, F) or not ops.gemm_uses_tensor_cores(X.shape[0], F, k):
        return None
    hs = getattr(layer, "_x_head", None)
    if hs is None or hs[0] != (id(X), k):
        hs = ((id(X), k), HeadSplit(X, k_head=k, top=None if plan is None else plan["top"]))
        layer._x_head = hs
    return hs[1]


def _x_product(layer, X, W, out, bias=None, act="identity"):
    """act(X.W + b) (S.dot of lasagne_layers.py:26,65): one SpMM, or dense head GEMM + sparse tail at Twitter scale."""
    hs = _head_split(layer, X, W.shape[1])
    if hs is None:
        return ops.spmm(X, W, bias=bias, act=act, out=out)
    return hs.product(W, out, bias=bias, act=act)


def _xt_product(layer, X, dZ, out):
    """dW = X^T.dZ (Dot.grad of lasagne_layers.py:26,65).  For large X the frequent terms are processed one
    document block at a time (sparse.BlockedRows) so that the gathered rows of dZ stay in L2, and the most
    frequent ones of all as a dense tensor-core product (sparse.HeadSplit)."""
    import os
    from .sparse import BlockedRows
    # B200 sweeps (profiles/r01_spmm_notes.md, profiles/r02_xt_sweep.jsonl): document blocks of 96 MB of dZ rows and
    # "heavy" = at least 16 non-zeros per block are the fastest pair at Twitter-World shape (39.0 ms vs 43.8 ms at 64 / 4)
    mb = float(os.environ.get("GCG_XT_BLOCK_MB", "96"))
    hf = int(os.environ.get("GCG_XT_HEAVY_FACTOR", "16"))
    hs = _head_split(layer, X, dZ.shape[1])
    Xs = X if hs is None else hs.tail
    red = getattr(layer, "_grad_reduce", None)      # multi-GPU: sum the pieces over ranks as they are finished
    plan = getattr(layer, "_x_plan", None)          # multi-GPU: the same blocking decisions on every rank
    big = _is_big(X, dZ.shape[1]) if plan is None else plan["big"]
    if mb <= 0 or not big:
        ops.spmm(Xs.T, dZ, out=out)
        if red is not None:
            red(out).wait()
    else:
        key = (id(X), dZ.shape[1], mb, hf)
        br = getattr(layer, "_xt_blocked", None)
        if br is None or br[0] != key:
            br = (key, BlockedRows(Xs.T, dZ.shape[1], block_mb=mb, heavy_factor=hf,
                                   heavy_ids=None if plan is None else plan["heavy_ids"]))
            layer._xt_blocked = br
        br[1].product(dZ, out, reduce=red)
    if hs is not None:
        hs.transpose_product_head(dZ, out, reduce=red)   # the head terms' rows are empty in the tail: plain placement
    return out


def _check_sparse(input):
    # lasagne_layers.py:22-24, 33-35, 61-63
    if not is_sparse(input):
        raise ValueError("Input for this layer must be sparse")


class SparseInputDenseLayer(DenseLayer):
    """act(X.W + b) for CSR X -- lasagne_layers.py:20-29."""

    def get_output_for(self, input, **kwargs):
        _check_sparse(input)
        X = as_csr(input, self.device)
        self._X = X
        n = X.shape[0]
        fused_act = "identity" if self.nonlinearity == "softmax" else self.nonlinearity
        out = _x_product(self, X, self.W, self._mat(("out", n), n, self.num_units), bias=self.b, act=fused_act)   # :26-29
        self._out = out
        return self._softmax_or(out, kwargs)

    def backward(self, grad_output, preact=False, **kwargs):
        """grad wrt the activation output (or the pre-activation when ``preact``).
        dW = X^T.dP (Dot.grad), db = colsum(dP).  X is an input: no grad is returned."""
        n = grad_output.shape[0]
        dP = grad_output if preact or self.nonlinearity in ("identity", "softmax") else \
            ops.act_bwd(grad_output, self._out, self.nonlinearity, out=self._mat(("dP", n), n, self.num_units))
        if self.b is not None:
            ops.colsum(dP, out=self._grad("b", self.b))
        _xt_product(self, self._X, dP, self._grad("W", self.W))
        return None


class SparseInputDropoutLayer(Layer):
    """Dropout on a sparse input -- lasagne_layers.py:31-52.  ``deterministic=True`` or
    p == 0 returns the input unchanged (:37-38).  The stochastic branch draws its mask
    from torch's CUDA generator; it cannot reproduce Theano's MRG stream, so parity
    configurations run with dropout off (as the reference driver does, tensormain.py:234)."""

    def __init__(self, incoming, p=0.5, rescale=True, **kwargs):
        super().__init__(incoming, **kwargs)
        self.p = float(p)
        self.rescale = rescale

    def get_output_for(self, input, deterministic=False, **kwargs):
        _check_sparse(input)
        if deterministic or self.p == 0:
            return input
        X = as_csr(input, self.device)
        retain = 1.0 - self.p
        keep = (torch.rand(X.data.shape, device=X.data.device) < retain).to(torch.float32)
        data = X.data * keep * ((1.0 / retain) if self.rescale else 1.0)      # :44-52
        host = None if X.host is None else (X.host[0], None, None)       # same pattern: the row offsets stay valid
        out = CSRMatrix(X.indptr, X.indices, data, X.shape, host=host, long_row_threshold=X.long_row_threshold)
        out.device_built = X.device_built
        return out


class TargetIndices:
    """A ``target_indices`` vector (mlpconv.py:179, tensormain.py:225-230) prepared once:
    device copy, the row-gathered A_hat[idx, :] (so only wanted rows are ever propagated)
    and the inverse map used by the deterministic scatter-add of the backward."""

    def __init__(self, idx, H: CSRMatrix):
        self.host = np.ascontiguousarray(np.asarray(idx), dtype=np.int32)
        self.n = len(self.host)
        self.device = H.device
        self.dev = torch.from_numpy(self.host).to(self.device)
        self.H = H
        self._Hsub = None
        self._pos = None

    @property
    def Hsub(self):
        if self._Hsub is None:
            self._Hsub = self.H.gather_rows(self.host)
        return self._Hsub

    @property
    def positions(self):
        if self._pos is None:
            self._pos = ops.scatter_positions(self.host, self.H.shape[0], self.device)
        return self._pos


class _ConvBase(DenseLayer):
    def __init__(self, incoming, H=None, **kwargs):
        super().__init__(incoming, **kwargs)
        self.H = None if H is None else as_csr(H, self.device)    # lasagne_layers.py:56,77 (not a param)
        self._ti_cache = {}

    def _operand(self, key, n_rows, n_cols):
        """buffer for a dense SpMM operand: in row-partitioned mode it is the local slab of the
        (shared, transient) all-gather buffer, so no staging copy is needed"""
        if hasattr(self.H, "operand"):
            return self.H.operand((id(self), key), n_cols)      # one gather buffer per layer and role: every
                                                                # operand of a step stays inspectable afterwards
        return self._mat(key, n_rows, n_cols)

    def _target(self, target_indices):
        if target_indices is None or isinstance(target_indices, TargetIndices):
            return target_indices
        key = (id(target_indices), len(target_indices))
        hit = self._ti_cache.get(key)
        if hit is None:
            hit = (target_indices, TargetIndices(target_indices, self.H))   # keep the array alive: id is the key
            self._ti_cache[key] = hit
        return hit[1]


class SparseConvolutionDenseLayer(_ConvBase):
    """act(H.(X.W) + b) for CSR X -- lasagne_layers.py:53-71 (GCN layer 1, mlpconv.py:205-209)."""

    def get_output_for(self, input, **kwargs):
        _check_sparse(input)                                               # :61-63
        X = as_csr(input, self.device)
        self._X = X
        N = X.shape[0]
        z = _x_product(self, X, self.W, self._operand("Z", N, self.num_units))  # :65
        out = ops.spmm(self.H, z, bias=self.b, act=self.nonlinearity,
                       out=self._mat("out", N, self.num_units))             # :67-71 (bias+act fused)
        self._out = out
        return out

    def backward(self, grad_output, preact=False, **kwargs):
        dP = grad_output if preact or self.nonlinearity == "identity" else \
            ops.act_bwd(grad_output, self._out, self.nonlinearity, out=self._operand("dP", *grad_output.shape))
        if self.b is not None:
            ops.colsum(dP, out=self._grad("b", self.b))
        dZ = ops.spmm(self.H, dP, out=self._operand("Z", *dP.shape))        # A_hat^T = A_hat; Z is dead: reuse
        _xt_product(self, self._X, dZ, self._grad("W", self.W))             # dW = X^T.dZ
        return None


class ConvolutionDenseLayer(_ConvBase):
    """act((H.(input.W) + b)[target_indices, :]) -- lasagne_layers.py:73-89.

    ``nonlinearity=softmax`` is the output layer of mlpconv.py:213-216.  Only the rows in
    ``target_indices`` are propagated (A_hat[idx,:].Z), which equals gathering afterwards.
    kwargs: ``target_indices`` (array or TargetIndices; None keeps every row),
    ``logits=True`` returns the pre-softmax rows (the training step feeds them to the
    fused softmax/CE head).

    ``propagate_first`` (ctor; "auto" = when num_units > num_inputs, e.g. 1024 regions from 600
    hidden units): evaluates (H.input).W instead of H.(input.W) -- the same product by
    associativity, but the sparse propagation then runs at the narrower width (and only for the
    target rows), which is the cheaper side of this HBM-bound layer.  Results differ from the
    reference order only by float32 rounding (tests/test_gpu_layers.py)."""

    def __init__(self, incoming, H=None, propagate_first="auto", **kwargs):
        super().__init__(incoming, H=H, **kwargs)
        self.propagate_first = (self.num_units > self.num_inputs) if propagate_first == "auto" else bool(propagate_first)

    def _forward_propagate_first(self, input, ti, kwargs):
        N = input.shape[0]
        Hm = self.H if ti is None else ti.Hsub
        n_out = N if ti is None else ti.n
        q = ops.spmm(Hm, input, out=self._mat(("Q", n_out), n_out, self.num_inputs))      # H[idx,:].input
        self._q = q
        self._s_q = self._split(("q", n_out), q, self.num_units)
        fused_act = "identity" if self.nonlinearity == "softmax" else self.nonlinearity
        out = ops.gemm(q, self.W, bias=self.b, act=fused_act, a_split=self._s_q,
                       chained=(self.nonlinearity == "softmax"),                          # logits cancel: short chains
                       out=self._mat(("out", n_out), n_out, self.num_units))               # (.).W + b
        self._out = out
        if self.nonlinearity == "softmax" and not kwargs.get("logits", False):
            probs = self._mat(("probs", n_out), n_out, self.num_units)
            ops.softmax_ce(out, probs=probs)
            return probs
        return out

    def _backward_propagate_first(self, grad_output, preact, input_mask, need_input_grad):
        ti = self._ti
        N = self._in.shape[0]
        if self.nonlinearity in ("softmax", "identity") or preact:
            dPr = grad_output
        else:
            dPr = ops.act_bwd(grad_output, self._out, self.nonlinearity, out=self._mat("dPr", *grad_output.shape))
        if self.b is not None:
            ops.colsum(dPr, out=self._grad("b", self.b))
        s_g = self._split(("dPr", dPr.shape[0]), dPr, self.num_inputs)
        ops.gemm(self._q, dPr, transA=True, out=self._grad("W", self.W), a_split=self._s_q, b_split=s_g)   # dW = Q^T.dP
        if not need_input_grad:
            return None
        dQ = ops.gemm(dPr, self.W, transB=True, a_split=s_g,
                      out=self._mat(("dQ", dPr.shape[0]), dPr.shape[0], self.num_inputs))
        if ti is not None:
            ptr, pos = ti.positions
            S = ops.scatter_rows(dQ, ptr, pos, N, out=self._operand("dP", N, self.num_inputs))
        else:
            S = dQ
        dIn = ops.spmm(self.H, S, out=self._mat("dIn", N, self.num_inputs))                # H^T = H
        if input_mask is not None:
            ops.act_bwd(dIn, input_mask[0], input_mask[1], out=dIn)
        return dIn

    def get_output_for(self, input, **kwargs):
        ti = self._target(kwargs.get("target_indices"))                    # :81
        N = input.shape[0]
        self._in = input
        self._ti = ti
        if self.propagate_first:
            return self._forward_propagate_first(input, ti, kwargs)
        self._s_in = self._split("in", input, self.num_units)
        z = ops.gemm(input, self.W, out=self._operand("Z", N, self.num_units), a_split=self._s_in,
                     chained=_FWD_CHAINED or self.nonlinearity == "softmax")                # :82
        Hm = self.H if ti is None else ti.Hsub
        n_out = N if ti is None else ti.n
        fused_act = "identity" if self.nonlinearity == "softmax" else self.nonlinearity
        out = ops.spmm(Hm, z, bias=self.b, act=fused_act,
                       out=self._mat(("out", n_out), n_out, self.num_units))  # :84-88
        self._out = out
        if self.nonlinearity == "softmax" and not kwargs.get("logits", False):
            probs = self._mat(("probs", n_out), n_out, self.num_units)
            ops.softmax_ce(out, probs=probs)                               # :89
            return probs
        return out

    def backward(self, grad_output, preact=False, input_mask=None, need_input_grad=True, **kwargs):
        """``grad_output``: grad wrt the layer output rows (wrt the LOGITS for a softmax
        layer -- the head produces it fused with the loss).  ``input_mask=(A_prev, act)``
        fuses the previous layer's act' into the dH product."""
        if self.propagate_first:
            return self._backward_propagate_first(grad_output, preact, input_mask, need_input_grad)
        ti = self._ti
        N = self._in.shape[0]
        if self.nonlinearity in ("softmax", "identity") or preact:
            dPr = grad_output
        else:
            dPr = ops.act_bwd(grad_output, self._out, self.nonlinearity,
                              out=self._mat("dPr", *grad_output.shape))
        if ti is not None:
            ptr, pos = ti.positions
            dP = ops.scatter_rows(dPr, ptr, pos, N, out=self._operand("dP", N, self.num_units))  # grad of :88
        else:
            dP = dPr
        if self.b is not None:
            ops.colsum(dP, out=self._grad("b", self.b))
        dZ = ops.spmm(self.H, dP, out=self._operand("Z", N, self.num_units))     # A_hat^T.dP
        s_dz = self._split("dZ", dZ, self.num_inputs)
        ops.gemm(self._in, dZ, transA=True, out=self._grad("W", self.W),
                 a_split=self._s_in, b_split=s_dz)                           # dW = H_in^T.dZ
        if not need_input_grad:
            return None
        mk, mact = (None, "identity") if input_mask is None else input_mask
        return ops.gemm(dZ, self.W, transB=True, mask=mk, mask_act=mact, a_split=s_dz,
                        out=self._mat("dIn", N, self.num_inputs))            # dH = dZ.W^T


class HighwayConvolutionDenseLayer(ConvolutionDenseLayer):
    """Gated conv layer (north_star; not in the reference):
        H' = act(A_hat.(H.W) + b);  g = sigmoid(H.W_g + b_g);  out = g*H' + (1-g)*H
    in-dim must equal out-dim.  The bias/act/mix epilogue is fused into the SpMM, so H'
    only reaches HBM in training mode (``train=True``), where the backward needs it."""

    def __init__(self, incoming, H=None, Wg=None, bg=Constant(0.0), rng=None, **kwargs):
        super().__init__(incoming, H=H, rng=rng, **kwargs)
        self.propagate_first = False
        assert self.num_inputs == self.num_units, "highway gate needs in-dim == out-dim"
        assert self.nonlinearity != "softmax"
        Wg = GlorotUniform() if Wg is None else Wg
        self.Wg = _to_param(Wg, (self.num_inputs, self.num_units), self.device, rng)
        self.bg = _to_param(bg, (self.num_units,), self.device, rng)
        self.params.append(("Wg", self.Wg, dict(trainable=True, regularizable=True)))
        self.params.append(("bg", self.bg, dict(trainable=True, regularizable=False)))

    def get_output_for(self, input, **kwargs):
        assert kwargs.get("target_indices") is None, "a gated layer keeps all rows"
        N, h = input.shape[0], self.num_units
        self._in = input
        self._s_in = self._split("in", input, h)                  # one hi/lo split feeds four GEMMs (fwd 2, bwd 2)
        z = ops.gemm(input, self.W, out=self._operand("Z", N, h), a_split=self._s_in, chained=_FWD_CHAINED)
        g = ops.gemm(input, self.Wg, bias=self.bg, act="sigmoid", out=self._mat("g", N, h), a_split=self._s_in)
        conv = self._mat("Hc", N, h) if kwargs.get("train", False) else None
        out = ops.spmm(self.H, z, bias=self.b, act=self.nonlinearity, gate=g, carry=input, conv_out=conv,
                       out=self._mat("out", N, h))
        self._g, self._Hc, self._out = g, conv, out
        return out

    def backward(self, grad_output, input_mask=None, **kwargs):
        assert self._Hc is not None, "forward must run with train=True before backward"
        N, h = grad_output.shape
        dP, dG, dIn = ops.highway_bwd(grad_output, self._g, self._Hc, self._in, self.nonlinearity,
                                      dP=self._operand("dP", N, h), dGpre=self._mat("dG", N, h),
                                      dHin=self._mat("dIn", N, h))
        ops.colsum(dP, out=self._grad("b", self.b))
        ops.colsum(dG, out=self._grad("bg", self.bg))
        dZ = ops.spmm(self.H, dP, out=self._operand("Z", N, h))
        s_dz, s_dg = self._split("dZ", dZ, h), self._split("dG", dG, h)
        ops.gemm(self._in, dZ, transA=True, out=self._grad("W", self.W), a_split=self._s_in, b_split=s_dz)
        ops.gemm(self._in, dG, transA=True, out=self._grad("Wg", self.Wg), a_split=self._s_in, b_split=s_dg)
        ops.gemm(dZ, self.W, transB=True, beta=1.0, out=dIn, a_split=s_dz)
        mk, mact = (None, "identity") if input_mask is None else input_mask
        ops.gemm(dG, self.Wg, transB=True, beta=1.0, mask=mk, mask_act=mact, out=dIn, a_split=s_dg)
        return dIn


def GraphConvLayer(incoming, H=None, sparse_input=False, highway=False, **kwargs):
    """north_star's name for the graph-conv layer: picks the reference class."""
    if sparse_input:
        return SparseConvolutionDenseLayer(incoming, H=H, **kwargs)
    if highway:
        return HighwayConvolutionDenseLayer(incoming, H=H, **kwargs)
    return ConvolutionDenseLayer(incoming, H=H, **kwargs)


# --------------------------------------------------------------------------- #
# lasagne.layers helper functions used by MLPCONV (mlpconv.py:222,226,262,301,312)
# --------------------------------------------------------------------------- #


def get_all_layers(layer):
    chain = []
    while layer is not None:
        chain.append(layer)
        layer = layer.input_layer
    return chain[::-1]


def get_output(layer, inputs=None, **kwargs):
    """lasagne.layers.get_output: run the chain; every layer sees the same kwargs."""
    x = inputs
    for ly in get_all_layers(layer):
        if isinstance(ly, InputLayer):
            x = ly.input_var if x is None else x
            continue
        x = ly.get_output_for(x, **kwargs)
    return x


def get_all_params(layer, **tags):
    out = []
    for ly in get_all_layers(layer):
        out += ly.get_params(**tags)
    return out


def get_all_param_values(layer):
    return [p.detach().cpu().numpy().copy() for p in get_all_params(layer)]


def set_all_param_values(layer, values):
    params = get_all_params(layer)
    if len(params) != len(values):
        raise ValueError("mismatch: got %d values to set %d parameters" % (len(values), len(params)))
    for p, v in zip(params, values):
        v = np.asarray(v, dtype=np.float32)
        if tuple(v.shape) != tuple(p.shape):
            raise ValueError("mismatch: parameter has shape %r but value to set has shape %r"
                             % (tuple(p.shape), v.shape))
        p.copy_(torch.from_numpy(np.ascontiguousarray(v)))
