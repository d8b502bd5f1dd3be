"""float32 NumPy/scipy restatement of the reference GCN hot path (TEST INFRASTRUCTURE).

Each function cites the reference file:line (under /root/reference) it follows.
The reference executes these through Theano/Lasagne on CPU; Theano's sparse
``S.dot`` dispatches to scipy's ``csr @ dense`` (sparsetools ``csr_matvecs``:
per row, per non-zero in CSR order, ``y += a * x`` with separately rounded
multiply and add) and ``T.dot`` to BLAS sgemm -- exactly the calls used here.

Parity status: the reference ships no tests / golden vectors and cannot run here as it is
(Python 2 + Theano + Lasagne), so fixtures were produced by EXECUTING ITS SOURCE under stub
theano / lasagne modules (scripts and outputs under tests/golden/):
  * forward layers, A_hat builder, geo_eval: pinned bit for bit (tests/test_layers_golden.py);
  * loss / regulariser composition, Adam, the fit loop, predict: pinned to a run of the reference's own
    MLPCONV.fit (tests/test_fit_golden.py; tolerances of float32 summation order);
  * hand-written backward of the >2-layer / gated extension: cross-checked against torch-CPU float64
    autograd (tests/test_oracle.py).  The highway gate is not in the reference; its spec is the formula in
    BASELINE.json ``north_star`` (see ``highway_mix``): parity unpinned.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

F32 = np.float32

# --------------------------------------------------------------------------- #
# A_hat construction                                                           #
# --------------------------------------------------------------------------- #


def build_ahat(adj, dtype="float32"):
    """A_hat = D^-1/2 (A with unit diagonal) D^-1/2, float64 then cast.

    Follows tensormain.py:170-180 (duplicated at main.py:513-522,
    tensormain.py:97-106) and the cast at tensormain.py:221:
      adj.setdiag(1); diags = adj.sum(axis=1); diags_sqrt = 1/sqrt(diags);
      inf -> 0; H = D * adj * D; H.astype(dtype).
    ``adj`` is the (binary, symmetric) user-user adjacency as any scipy sparse
    matrix; it is not modified.
    """
    import warnings
    adj = sp.csr_matrix(adj, dtype=np.float64, copy=True)   # nx.adjacency_matrix gave the reference a CSR matrix (:170)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", sp.SparseEfficiencyWarning)
        adj.setdiag(1)                                  # tensormain.py:172
    n, m = adj.shape
    diags = np.asarray(adj.sum(axis=1)).flatten()       # :174
    with np.errstate(divide="ignore"):
        diags_sqrt = 1.0 / np.sqrt(diags)               # :175-176
    diags_sqrt[np.isinf(diags_sqrt)] = 0                # :177
    d = sp.spdiags(diags_sqrt, [0], m, n, format="csr")  # :178
    h = d * adj * d                                     # :179
    h = sp.csr_matrix(h.astype(dtype))                  # :180, :221
    h.sort_indices()
    return h


# --------------------------------------------------------------------------- #
# Non-linearities (Lasagne names; mlpconv.py:186-193)                          #
# --------------------------------------------------------------------------- #


def _softmax_rows(x):
    e = x - x.max(axis=1, keepdims=True)
    np.exp(e, out=e)
    e /= e.sum(axis=1, keepdims=True, dtype=F32)
    return e


def _sigmoid(x):
    return (F32(1) / (F32(1) + np.exp(-x, dtype=F32))).astype(F32)


ACTIVATIONS = {
    "identity": lambda x: x,
    "linear": lambda x: x,
    "rectify": lambda x: np.maximum(x, F32(0)),
    "tanh": lambda x: np.tanh(x, dtype=F32),
    "sigmoid": _sigmoid,
    "softmax": _softmax_rows,
}


def _act(name):
    if callable(name):
        return name
    return ACTIVATIONS[name]


# --------------------------------------------------------------------------- #
# Layer forwards                                                               #
# --------------------------------------------------------------------------- #


def sparse_input_dense(X, W, b, nonlinearity="rectify"):
    """SparseInputDenseLayer.get_output_for -- lasagne_layers.py:20-29."""
    if not sp.issparse(X):
        raise ValueError("Input for this layer must be sparse")   # :22-24
    activation = np.asarray(X @ W, dtype=F32)                     # :26 S.dot
    if b is not None:
        activation = activation + b[None, :]                      # :27-28
    return _act(nonlinearity)(activation)                         # :29


def sparse_convolution_dense(X, W, b, H, nonlinearity="rectify"):
    """SparseConvolutionDenseLayer.get_output_for -- lasagne_layers.py:60-71."""
    if not sp.issparse(X):
        raise ValueError("Input for this layer must be sparse")   # :61-63
    activation = np.asarray(X @ W, dtype=F32)                     # :65
    activation = np.asarray(H @ activation, dtype=F32)            # :67 the convolution
    if b is not None:
        activation = activation + b[None, :]                      # :69-70
    return _act(nonlinearity)(activation)                         # :71


def convolution_dense(inp, W, b, H, target_indices=None, nonlinearity="rectify"):
    """ConvolutionDenseLayer.get_output_for -- lasagne_layers.py:80-89.

    Note the order: propagate, then bias, then row gather, then non-linearity.
    ``target_indices=None`` keeps all rows (used for hidden layers of the
    >2-layer extension; the reference always passes indices, mlpconv.py:222,226).
    """
    activation = np.dot(inp, W).astype(F32)                       # :82 T.dot
    activation = np.asarray(H @ activation, dtype=F32)            # :84
    if b is not None:
        activation = activation + b[None, :]                      # :86-87
    if target_indices is not None:
        activation = activation[target_indices, :]                # :88
    return _act(nonlinearity)(activation)                         # :89


def highway_mix(h_in, h_conv, Wg, bg):
    """Highway gate -- NOT in the reference; formula from BASELINE.json north_star:
    g = sigmoid(H.W_g + b_g);  out = g * H' + (1 - g) * H.   (parity unpinned)
    """
    g = _sigmoid(np.dot(h_in, Wg).astype(F32) + bg[None, :])
    return (g * h_conv + (F32(1) - g) * h_in).astype(F32), g


# --------------------------------------------------------------------------- #
# Parameter init                                                               #
# --------------------------------------------------------------------------- #


def glorot_uniform(rng, fan_in, fan_out):
    """lasagne.init.GlorotUniform (mlpconv.py:208): U(-a, a), a = sqrt(6/(in+out))."""
    a = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-a, a, size=(fan_in, fan_out)).astype(F32)


def init_params(rng, in_size, hidden, out_size, n_layers=2, highway=False):
    """Parameter list in Lasagne ``get_all_param_values`` order.

    Reference net (mlpconv.py:196-217): [W1, b1, W2, b2].  Extension: hidden
    conv layers 2..L-1 contribute [W, b] (+ [Wg, bg] when gated).
    Biases start at 0 (DenseLayer default); gate bias 0 (unpinned choice).
    """
    params = [glorot_uniform(rng, in_size, hidden), np.zeros(hidden, F32)]
    for _ in range(n_layers - 2):
        params += [glorot_uniform(rng, hidden, hidden), np.zeros(hidden, F32)]
        if highway:
            params += [glorot_uniform(rng, hidden, hidden), np.zeros(hidden, F32)]
    params += [glorot_uniform(rng, hidden, out_size), np.zeros(out_size, F32)]
    return params


# --------------------------------------------------------------------------- #
# The network: forward, loss, hand-written backward                            #
# --------------------------------------------------------------------------- #


class GCNOracle:
    """L-layer GCN exactly as MLPCONV.fit builds it for L=2 (mlpconv.py:196-245).

    layer 1      : SparseConvolutionDenseLayer(V -> h, act)          mlpconv.py:205-209
    layers 2..L-1: ConvolutionDenseLayer(h -> h, act) [+ highway]    extension (north_star)
    layer L      : ConvolutionDenseLayer(h -> C, softmax) on idx     mlpconv.py:213-216
    loss         : mean CE + elastic net on the W's                  mlpconv.py:228-245
    regul_coefs unpack order is (out, hid) (mlpconv.py:237); every non-output W
    (gate weights included -- unpinned) uses the ``hid`` coefficient.
    """

    def __init__(self, X, A, n_layers=2, highway=False, regul_coefs=(5e-5, 5e-5),
                 nonlinearity="rectify"):
        self.X = sp.csr_matrix(X, dtype=F32)
        self.A = sp.csr_matrix(A, dtype=F32)
        self.XT = sp.csr_matrix(self.X.T)       # Dot.grad uses x.T (CSC view of X)
        self.L = n_layers
        self.highway = highway
        self.c_out, self.c_hid = (F32(c) for c in regul_coefs)
        self.act = nonlinearity
        assert nonlinearity in ("rectify", "tanh", "identity")

    # ---- helpers ---------------------------------------------------------- #
    def unpack(self, params):
        """-> list of per-layer dicts."""
        it = iter(params)
        layers = [dict(W=next(it), b=next(it))]
        for _ in range(self.L - 2):
            d = dict(W=next(it), b=next(it))
            if self.highway:
                d.update(Wg=next(it), bg=next(it))
            layers.append(d)
        layers.append(dict(W=next(it), b=next(it)))
        return layers

    def _dact(self, a_out):
        """act'(p) expressed from the activation output."""
        if self.act == "rectify":
            return (a_out > 0).astype(F32)
        if self.act == "tanh":
            return (F32(1) - a_out * a_out).astype(F32)
        return np.ones_like(a_out)

    def reg_loss(self, layers):
        """mlpconv.py:235-244: 0.5*c*(l1+l2) per regularised W; biases excluded."""
        tot = F32(0)
        for i, ly in enumerate(layers):
            c = self.c_out if i == len(layers) - 1 else self.c_hid
            for k in ("W", "Wg"):
                if k in ly:
                    w = ly[k]
                    tot = tot + F32(0.5) * c * (np.abs(w).sum(dtype=F32) + (w * w).sum(dtype=F32))
        return F32(tot)

    # ---- forward ---------------------------------------------------------- #
    def forward(self, params, idx, lean=False):
        """Temporaries are released / reused in place where that does not change a single rounding: at
        Twitter-World size every [N, F] float32 array is 3.4-5.7 GB and the epoch must fit the host."""
        layers = self.unpack(params)
        A, act = self.A, self.act
        c = {"layers": layers}

        def activate(p):                                                 # in place; p is not needed afterwards
            if act == "rectify":
                return np.maximum(p, F32(0), out=p)
            if act == "tanh":
                return np.tanh(p, out=p)
            return p
        z = np.asarray(self.X @ layers[0]["W"], dtype=F32)               # lasagne_layers.py:65
        p = np.asarray(A @ z, dtype=F32)                                 # :67
        p += layers[0]["b"][None, :]                                     # :69-70
        h = activate(p)                                                  # :71
        c["Z1"], c["A"] = (None if lean else z), [h]
        del z
        c["Hc"], c["g"] = [None], [None]
        for ly in layers[1:-1]:
            z = np.asarray(np.dot(h, ly["W"]), dtype=F32)                # :82
            p = np.asarray(A @ z, dtype=F32)                             # :84
            del z
            p += ly["b"][None, :]                                        # :86-87
            hc = activate(p)
            if self.highway:
                out, g = highway_mix(h, hc, ly["Wg"], ly["bg"])
            else:
                out, g = hc, None
            c["Hc"].append(hc)
            c["g"].append(g)
            c["A"].append(out)
            h = out
        ly = layers[-1]
        z = np.asarray(np.dot(h, ly["W"]), dtype=F32)                    # :82
        p = np.asarray(A @ z, dtype=F32)                                 # :84
        del z
        p += ly["b"][None, :]                                            # :86-87
        logits = p[idx, :]                                               # :88
        del p
        c["logits"] = logits
        c["probs"] = _softmax_rows(logits)                               # :89, mlpconv.py:216
        return c

    def predict_proba(self, params, idx):
        return self.forward(params, idx)["probs"]

    def predict(self, params, idx):
        return self.forward(params, idx)["probs"].argmax(-1)             # mlpconv.py:223

    def loss_acc(self, params, idx, y, cache=None):
        """mlpconv.py:227-233, 252: mean CE + reg; acc = mean(argmax == y)."""
        c = cache or self.forward(params, idx)
        lg = c["logits"]
        m = lg.max(axis=1, keepdims=True)
        e = lg - m
        np.exp(e, out=e)
        lse = (m[:, 0] + np.log(e.sum(axis=1, dtype=F32))).astype(F32)
        del e
        ce = (lse - lg[np.arange(len(y)), y]).astype(F32)
        loss = F32(ce.mean(dtype=F32)) + self.reg_loss(c["layers"])
        acc = float(np.mean(c["probs"].argmax(-1) == y))
        return F32(loss), acc

    # ---- backward (validated vs torch autograd) ------- #
    def backward(self, cache, idx, y):
        layers, A = cache["layers"], self.A
        n_idx = len(idx)
        N = A.shape[0]
        G = cache["probs"] if cache.get("keep", None) == () else cache["probs"].copy()   # lean: reuse the buffer
        G[np.arange(n_idx), y] -= F32(1)
        G /= F32(n_idx)
        C = G.shape[1]
        dP = np.zeros((N, C), F32)
        idx = np.asarray(idx)
        if len(np.unique(idx)) == n_idx:
            dP[idx] = G                          # no duplicates: the scatter-add is a placement
        else:
            np.add.at(dP, idx, G)                # scatter-ADD: duplicates accumulate
        grads = [None] * len(layers)
        acts = cache["A"]
        # output layer
        ly = layers[-1]
        h_in = acts[-1]
        db = dP.sum(axis=0, dtype=F32)
        dZ = np.asarray(A @ dP, dtype=F32)       # A^T = A
        if "dP_out" not in cache.get("keep", ("dP_out",)):
            del dP, G
            cache["probs"] = None
        dW = np.dot(h_in.T, dZ).astype(F32) + F32(0.5) * self.c_out * (np.sign(ly["W"]) + F32(2) * ly["W"])
        dH = np.asarray(np.dot(dZ, ly["W"].T), dtype=F32)
        del dZ
        grads[-1] = dict(W=dW.astype(F32), b=db)
        # hidden conv layers L-1 .. 2
        for li in range(len(layers) - 2, 0, -1):
            ly = layers[li]
            h_in, hc, g = acts[li - 1], cache["Hc"][li], cache["g"][li]
            gr = {}
            if self.highway:
                dHc = g * dH
                dg = dH * (hc - h_in)
                dGpre = (dg * g * (F32(1) - g)).astype(F32)
                dH_in = (F32(1) - g) * dH
                gr["bg"] = dGpre.sum(axis=0, dtype=F32)
                gr["Wg"] = (np.dot(h_in.T, dGpre).astype(F32)
                            + F32(0.5) * self.c_hid * (np.sign(ly["Wg"]) + F32(2) * ly["Wg"])).astype(F32)
                dH_in = dH_in + np.dot(dGpre, ly["Wg"].T).astype(F32)
            else:
                dHc, dH_in = dH, F32(0)
            dPl = (dHc * self._dact(hc)).astype(F32)
            gr["b"] = dPl.sum(axis=0, dtype=F32)
            dZ = np.asarray(A @ dPl, dtype=F32)
            gr["W"] = (np.dot(h_in.T, dZ).astype(F32)
                       + F32(0.5) * self.c_hid * (np.sign(ly["W"]) + F32(2) * ly["W"])).astype(F32)
            dH = (dH_in + np.dot(dZ, ly["W"].T).astype(F32)).astype(F32)
            grads[li] = gr
        # layer 1 (sparse input)
        ly = layers[0]
        dP1 = (dH * self._dact(acts[0])).astype(F32)
        db1 = dP1.sum(axis=0, dtype=F32)
        dZ1 = np.asarray(A @ dP1, dtype=F32)
        dW1 = (np.asarray(self.XT @ dZ1, dtype=F32)
               + F32(0.5) * self.c_hid * (np.sign(ly["W"]) + F32(2) * ly["W"])).astype(F32)
        grads[0] = dict(W=dW1, b=db1)
        cache["dZ1"], cache["dP1"] = dZ1, dP1
        if "dP_out" in cache.get("keep", ("dP_out",)):
            cache["dP_out"] = dP
        # flatten in parameter order
        flat = []
        for gr in grads:
            flat += [gr["W"], gr["b"]]
            if "Wg" in gr:
                flat += [gr["Wg"], gr["bg"]]
        return flat

    def loss_and_grads(self, params, idx, y, lean=False):
        """``lean``: drop the intermediates only tests look at (memory at Twitter-World size)."""
        c = self.forward(params, idx, lean=lean)
        if lean:
            c["keep"] = ()
        loss, acc = self.loss_acc(params, idx, y, cache=c)
        if lean:
            c["logits"] = None
        return loss, acc, self.backward(c, idx, y), c


# --------------------------------------------------------------------------- #
# Optimiser                                                                    #
# --------------------------------------------------------------------------- #


class AdamState:
    def __init__(self, params):
        self.t = F32(0)
        self.m = [np.zeros_like(p) for p in params]
        self.v = [np.zeros_like(p) for p in params]


def adam_step(params, grads, state, lr=4e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    """lasagne.updates.adam as called at mlpconv.py:263 (restated literally):
    t <- t+1; a_t = lr*sqrt(1-b2^t)/(1-b1^t); m <- b1 m + (1-b1) g;
    v <- b2 v + (1-b2) g^2; p <- p - a_t*m/(sqrt(v)+eps).  All float32.
    Updates ``params`` and ``state`` in place.
    """
    lr, b1, b2, eps, one = F32(lr), F32(beta1), F32(beta2), F32(eps), F32(1)
    state.t = F32(state.t + one)
    a_t = F32(lr * np.sqrt(one - b2 ** state.t, dtype=F32) / (one - b1 ** state.t))
    for p, g, m, v in zip(params, grads, state.m, state.v):
        m[...] = b1 * m + (one - b1) * g
        v[...] = b2 * v + (one - b2) * g * g
        p[...] = p - a_t * m / (np.sqrt(v, dtype=F32) + eps)
    return a_t


def train_epochs(net, params, idx, y, n_epochs, lr=4e-3):
    """The hot loop mlpconv.py:293-295: one f_train (fwd+bwd+Adam) per epoch.
    Returns the per-epoch (loss, acc) the reference logs (mlpconv.py:306)."""
    state = AdamState(params)
    hist = []
    for _ in range(n_epochs):
        loss, acc, grads, _ = net.loss_and_grads(params, idx, y)
        adam_step(params, grads, state, lr=lr)
        hist.append((float(loss), acc))
    return hist


def fit_loop(net, params, train_idx, y_train, dev_idx, y_dev, n_epochs, early_stopping_max_down=100000, lr=4e-3):
    """The epoch loop of MLPCONV.fit, mlpconv.py:288-317: one f_train per epoch; every 10th epoch f_val on dev, a
    strictly better dev loss snapshots the parameters (after that epoch's update) and resets the counter, anything
    else increments it; stop when the counter EXCEEDS early_stopping_max_down; finally restore the best snapshot
    and validate once more.  ``params`` is updated in place and ends as the best snapshot.
    Returns (train history [(loss, acc)], val history [(loss, acc)] incl. the final one, best parameter list)."""
    state = AdamState(params)
    train_hist, val_hist = [], []
    best, best_val_loss, n_down = None, float("inf"), 0
    for n in range(n_epochs):
        loss, acc, grads, _ = net.loss_and_grads(params, train_idx, y_train)       # :295 (loss before the update)
        adam_step(params, grads, state, lr=lr)
        train_hist.append((float(loss), acc))
        if n % 10 == 0:                                                           # :296
            l_val, a_val = net.loss_acc(params, dev_idx, y_dev)                   # :297
            val_hist.append((float(l_val), a_val))
            if l_val < best_val_loss:                                             # :298-302
                best_val_loss, best, n_down = l_val, [p.copy() for p in params], 0
            else:
                n_down += 1                                                       # :305
            if n_down > early_stopping_max_down:                                  # :307-309
                break
    for p, b in zip(params, best):                                                # :314
        p[...] = b
    l_val, a_val = net.loss_acc(params, dev_idx, y_dev)                           # :317
    val_hist.append((float(l_val), a_val))
    return train_hist, val_hist, best
