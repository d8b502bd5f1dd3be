"""float32 NumPy/scipy restatement of the reference's input-smoothing MLP path (TEST INFRASTRUCTURE).

SURVEY.md section 8(f) row 2: the one-shot smoothing ``X_conv = H * X`` (main.py:528-534,
tensormain.py:112-118) followed by the minibatch ``MLP`` whose first layer is the
``SparseInputDenseLayer`` of the hot path (mlp.py:36-45, 121-314).

Parity status: the SpGEMM is PINNED -- it *is* scipy's ``csr_matmat`` (the routine the reference
calls through ``H * X``), called here directly, and tests compare the CUDA product with it bit for
bit.  The MLP training loop is UNPINNED like the GCN (Theano/Lasagne cannot run here); the
restatement follows mlp.py line by line and its backward is checked against torch-CPU autograd in
tests/test_oracle.py.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .gcn_oracle import F32, AdamState, _act, _softmax_rows, adam_step, glorot_uniform


def smooth_features(H, X):
    """``X_conv = H * X; X_conv = X_conv.tocsr().astype('float32')`` -- main.py:528-530.
    H stays float64 in that driver (main.py:513-522 never casts it), X is float32 (data.py:390):
    scipy upcasts to float64, accumulates in csr_matmat order and the result is rounded once."""
    xc = sp.csr_matrix(H) * sp.csr_matrix(X)
    return xc.tocsr().astype("float32")


def iterate_minibatches(n, batchsize, rng, shuffle=True):
    """Index version of mlp.py:81-91: only FULL batches are yielded (range stops at n - batchsize + 1)."""
    indices = np.arange(n)
    if shuffle:
        rng.shuffle(indices)                                         # mlp.py:84-85
    for start in range(0, n - batchsize + 1, batchsize):
        yield indices[start:start + batchsize]


def init_params(rng, in_size, hidden, out_size, add_hidden=True):
    """[W1, b1, W2, b2] (mlp.py:171-185) or [W, b] without the hidden layer (mlp.py:193-195)."""
    if not add_hidden:
        return [glorot_uniform(rng, in_size, out_size), np.zeros(out_size, F32)]
    return [glorot_uniform(rng, in_size, hidden), np.zeros(hidden, F32),
            glorot_uniform(rng, hidden, out_size), np.zeros(out_size, F32)]


class MLPOracle:
    """The network MLP.fit builds (mlp.py:152-232), dropout off."""

    def __init__(self, regul_coefs=(5e-5, 5e-5), nonlinearity="rectify", add_hidden=True):
        self.c_out, self.c_hid = (F32(c) for c in regul_coefs)       # mlp.py:222 (out, hid)
        self.nonlinearity = nonlinearity
        self.add_hidden = add_hidden

    def forward(self, params, X):
        if self.add_hidden:
            W1, b1, W2, b2 = params
            z = X @ W1 if sp.issparse(X) else np.dot(X, W1)
            a1 = _act(self.nonlinearity)(np.asarray(z, dtype=F32) + b1[None, :])     # mlp.py:40-45 / DenseLayer
            logits = np.dot(a1, W2).astype(F32) + b2[None, :]                        # mlp.py:183-185
            return logits, a1
        W, b = params
        z = X @ W if sp.issparse(X) else np.dot(X, W)
        return np.asarray(z, dtype=F32) + b[None, :], None

    def reg_loss(self, params):
        """mlp.py:220-229: 0.5*c*(l1 + l2) for the W of each layer, biases excluded."""
        pen = lambda W, c: F32(0.5) * c * (np.abs(W).sum(dtype=F32) + (W * W).sum(dtype=F32))
        if self.add_hidden:
            return F32(pen(params[2], self.c_out) + pen(params[0], self.c_hid))
        return F32(pen(params[0], self.c_out))

    def loss_acc(self, params, X, y):
        logits, _ = self.forward(params, X)
        S = _softmax_rows(logits)
        n = len(y)
        ce = -np.log(S[np.arange(n), y], dtype=F32).mean(dtype=F32)  # mlp.py:209-210,214-216
        acc = float((S.argmax(-1) == y).mean())                      # mlp.py:236-237
        return F32(ce + self.reg_loss(params)), acc

    def predict_proba(self, params, X):
        return _softmax_rows(self.forward(params, X)[0])

    def loss_and_grads(self, params, X, y):
        logits, a1 = self.forward(params, X)
        S = _softmax_rows(logits)
        n = len(y)
        loss = F32(-np.log(S[np.arange(n), y], dtype=F32).mean(dtype=F32) + self.reg_loss(params))
        acc = float((S.argmax(-1) == y).mean())
        G = S.copy()
        G[np.arange(n), y] -= F32(1)
        G /= F32(n)
        reg_grad = lambda W, c: F32(0.5) * c * (np.sign(W) + F32(2) * W)
        if not self.add_hidden:
            W, b = params
            dW = np.asarray(X.T @ G, dtype=F32) + reg_grad(W, self.c_out)
            return loss, acc, [dW.astype(F32), G.sum(axis=0, dtype=F32)]
        W1, b1, W2, b2 = params
        dW2 = np.dot(a1.T, G).astype(F32) + reg_grad(W2, self.c_out)
        db2 = G.sum(axis=0, dtype=F32)
        dA1 = np.dot(G, W2.T).astype(F32)
        if self.nonlinearity == "rectify":
            dP1 = dA1 * (a1 > 0)
        elif self.nonlinearity == "tanh":
            dP1 = dA1 * (F32(1) - a1 * a1)
        elif self.nonlinearity == "sigmoid":
            dP1 = dA1 * a1 * (F32(1) - a1)
        else:
            dP1 = dA1
        dP1 = dP1.astype(F32)
        xt = X.T @ dP1 if sp.issparse(X) else np.dot(X.T, dP1)       # Dot.grad of mlp.py:42
        dW1 = np.asarray(xt, dtype=F32) + reg_grad(W1, self.c_hid)
        db1 = dP1.sum(axis=0, dtype=F32)
        return loss, acc, [dW1.astype(F32), db1, dW2.astype(F32), db2]


def fit(net, params, X_train, Y_train, X_dev, Y_dev, n_epochs, batch_size, seed, lr=2e-3,
        early_stopping_max_down=100000):
    """The loop of mlp.py:261-283: shuffled full minibatches, Adam(lr=2e-3) per batch (mlp.py:244,269-271),
    dev evaluation per epoch, best parameters by dev ACCURACY (mlp.py:273-277), early stopping.
    The reference shuffles with NumPy's global generator; ``seed`` makes the same draw reproducible.
    Returns (history of per-batch (loss, acc), per-epoch (val_loss, val_acc), best params)."""
    rng = np.random.RandomState(seed)
    state = AdamState(params)
    Y_train = np.asarray(Y_train, dtype=np.int32)
    Y_dev = np.asarray(Y_dev, dtype=np.int32)
    steps, epochs = [], []
    best_acc, best_params, down = 0.0, None, 0
    for _ in range(n_epochs):
        for idx in iterate_minibatches(X_train.shape[0], batch_size, rng, shuffle=True):
            loss, acc, grads = net.loss_and_grads(params, X_train[idx], Y_train[idx])
            adam_step(params, grads, state, lr=lr)
            steps.append((float(loss), acc))
        l_val, acc_val = net.loss_acc(params, X_dev, Y_dev)
        epochs.append((float(l_val), acc_val))
        if acc_val > best_acc:
            best_acc, best_params, down = acc_val, [p.copy() for p in params], 0
        else:
            down += 1
        if down > early_stopping_max_down:
            break
    return steps, epochs, best_params
