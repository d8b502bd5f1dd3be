"""``MLP`` -- the minibatch classifier the reference trains on the smoothed features, on B200.

Same constructor keywords and ``fit / predict / predict_proba / accuracy / score / get_embedding``
signatures as /root/reference/mlp.py:121-314, so ``main.py:552-556`` runs unchanged:

    X_conv = smooth_features(H, X)            # main.py:528-530  (sparse.smooth_features, SpGEMM on the GPU)
    clf = MLP(n_epochs=200, batch_size=batch_size, ..., hidden_layer_size=hidden_size, drop_out=True, ...)
    clf.fit(X_train, Y_train, X_dev, Y_dev)

The first layer is the hot path's ``SparseInputDenseLayer`` (mlp.py:36-45 == lasagne_layers.py:20-29).
Minibatches (`inputs[excerpt]`, mlp.py:81-91) are sliced out of the device-resident CSR by
gcg_csr_gather_rows_device and transposed for dW = X_b^T.dP by gcg_csr_transpose_device: the training
data crosses PCIe once per fit, not once per batch.

Reference quirks kept: only FULL batches are used (mlp.py:86); the best parameters are chosen by dev
ACCURACY (mlp.py:273); the hidden-layer dropout is built but never connected to the output layer
(mlp.py:181-185 use ``l_hid1``, not ``self.l_hid1``), so only the input dropout is active.
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from . import lasagne_layers as L
from . import ops
from .sparse import CSRMatrix, RowBlockedCSR, as_csr, is_sparse

logger = logging.getLogger("graphconvgeo_b200")


class MLP:
    def __init__(self,
                 n_epochs=10,
                 batch_size=1000,
                 init_parameters=None,
                 complete_prob=False,
                 add_hidden=True,
                 regul_coefs=[5e-5, 5e-5],
                 save_results=False,
                 hidden_layer_size=None,
                 drop_out=False,
                 drop_out_coefs=[0.5, 0.5],
                 early_stopping_max_down=100000,
                 loss_name='log',
                 nonlinearity='rectify',
                 # ---- extensions ----
                 device='cuda',
                 seed=None,
                 learning_rate=2e-3):
        # mlp.py:136-148
        self.n_epochs = n_epochs
        self.batch_size = batch_size
        self.init_parameters = init_parameters
        self.complete_prob = complete_prob
        self.add_hidden = add_hidden
        self.regul_coefs = regul_coefs
        self.save_results = save_results
        self.hidden_layer_size = hidden_layer_size
        self.drop_out = drop_out
        self.drop_out_coefs = drop_out_coefs
        self.early_stopping_max_down = early_stopping_max_down
        self.loss_name = loss_name
        # the reference hard-codes 'rectify' (mlp.py:148); we honour the argument
        self.nonlinearity = nonlinearity if nonlinearity in ('rectify', 'sigmoid', 'tanh') else 'rectify'
        self.device = torch.device(device)
        self.seed = seed
        self.learning_rate = learning_rate
        if complete_prob:
            raise NotImplementedError("complete_prob=True (dense label distributions) is unused by the "
                                      "reference drivers (main.py:552) and not on the hot path")
        if loss_name != 'log':
            raise ValueError("only loss_name='log' works in the reference (mlp.py:213-215 reference undefined names)")

    # ------------------------------------------------------------------ data
    def _to_device(self, X):
        if is_sparse(X):
            if isinstance(X, (CSRMatrix, RowBlockedCSR)):
                return X
            return CSRMatrix.from_scipy(X, device=self.device, sort_indices=False)   # keep scipy's entry order
        t = torch.as_tensor(np.ascontiguousarray(X, dtype=np.float32)).to(self.device)
        buf = ops.alloc_mat(t.shape[0], t.shape[1], self.device)
        buf.copy_(t)
        return buf

    def _labels(self, y):
        return torch.from_numpy(np.ascontiguousarray(y, dtype=np.int32)).to(self.device)   # mlp.py:259

    # ----------------------------------------------------------------- build
    def _build(self, sparse_input, in_size, out_size):
        rng = np.random.RandomState(self.seed) if self.seed is not None else None
        drop_out_hid, drop_out_in = self.drop_out_coefs                     # mlp.py:132
        l_in = L.InputLayer(shape=(None, in_size), device=self.device)      # mlp.py:155-156
        first = l_in
        if self.drop_out:                                                   # mlp.py:167-168
            first = L.SparseInputDropoutLayer(l_in, p=drop_out_in) if sparse_input else \
                _DenseDropout(l_in, p=drop_out_in)
        first_cls = L.SparseInputDenseLayer if sparse_input else L.DenseLayer
        if self.add_hidden:
            self.l_hid1 = first_cls(first, num_units=self.hidden_layer_size, nonlinearity=self.nonlinearity,
                                    W=L.GlorotUniform(), rng=rng)           # mlp.py:171-180
            self.l_out = L.DenseLayer(self.l_hid1, num_units=out_size,
                                      nonlinearity=L.nonlinearities.softmax, rng=rng)   # mlp.py:183-185
        else:
            self.l_hid1 = None
            self.l_out = first_cls(first, num_units=out_size, nonlinearity=L.nonlinearities.softmax,
                                   rng=rng)                                 # mlp.py:187-195
        self.layers = [ly for ly in L.get_all_layers(self.l_out) if not isinstance(ly, L.InputLayer)]
        if self.init_parameters is not None:
            L.set_all_param_values(self.l_out, self.init_parameters)        # mlp.py:238-239
        regul_coef_out, regul_coef_hid = self.regul_coefs                   # mlp.py:222
        params, grads, reg = [], [], []
        for ly in self.layers:
            for name, t, tags in ly.params:
                if not tags.get("trainable"):
                    continue
                g = torch.zeros_like(t)
                ly.grads[name] = g
                params.append(t)
                grads.append(g)
                coef = regul_coef_out if ly is self.l_out else regul_coef_hid
                reg.append(coef if tags.get("regularizable") else 0.0)
        self.params, self.grads = params, grads
        self.adam = ops.Adam(params, grads, reg, lr=self.learning_rate, beta1=0.9, beta2=0.999, eps=1e-8)   # mlp.py:244
        self.elastic = ops.ElasticNet(params, reg)
        self._heads = {}

    def _head_buffers(self, n):
        hb = self._heads.get(n)
        if hb is None:
            d = self.device
            hb = dict(ce=torch.empty(n, dtype=torch.float32, device=d),
                      hit=torch.empty(n, dtype=torch.float32, device=d),
                      pred=torch.empty(n, dtype=torch.int64, device=d),
                      out=torch.zeros(2, dtype=torch.float32, device=d))
            self._heads[n] = hb
        return hb

    def _forward(self, x, train):
        for ly in self.layers:
            x = ly.get_output_for(x, logits=True, deterministic=not train)
        return x

    # -------------------------------------------------------------- functions
    def f_train(self, x_batch, y_batch):
        """One minibatch update (mlp.py:246, 269-270).  Returns the device buffers holding
        [mean CE, accuracy]; the elastic-net penalty of the pre-update parameters is in adam.reg_out."""
        logits = self._forward(x_batch, train=True)
        n, C = logits.shape
        hb = self._head_buffers(n)
        G = self.l_out._mat(("G", n), n, C)
        ops.softmax_ce(logits, y=y_batch, grad=G, ce=hb["ce"], hit=hb["hit"], denom=n)
        ops.sum_scaled(hb["ce"], 1.0 / n, out=hb["out"][0:1])
        ops.sum_scaled(hb["hit"], 1.0 / n, out=hb["out"][1:2])
        if self.add_hidden:
            hid = self.l_hid1
            dP = self.l_out.backward(G, input_mask=(hid._out, hid.nonlinearity))    # act' of the hidden layer fused
            hid.backward(dP, preact=True, need_input_grad=False)
        else:
            self.l_out.backward(G, need_input_grad=False)
        self.adam.step()
        self._train_hb = hb
        return hb

    def train_results(self):
        o = self._train_hb["out"].cpu().numpy()
        return float(np.float32(o[0]) + np.float32(self.adam.reg_out.item())), float(o[1])

    def f_val(self, X, y):
        """[eval_loss, eval_acc] (mlp.py:247): deterministic forward, CE + the same penalty."""
        logits = self._forward(X, train=False)
        n = logits.shape[0]
        hb = self._head_buffers(n)
        ops.softmax_ce(logits, y=y, ce=hb["ce"], hit=hb["hit"], denom=n)
        ops.sum_scaled(hb["ce"], 1.0 / n, out=hb["out"][0:1])
        ops.sum_scaled(hb["hit"], 1.0 / n, out=hb["out"][1:2])
        reg = self.elastic()
        o = hb["out"].cpu().numpy()
        return float(np.float32(o[0]) + np.float32(reg.item())), float(o[1])

    def f_predict(self, X):
        logits = self._forward(X, train=False)
        hb = self._head_buffers(logits.shape[0])
        ops.softmax_ce(logits, pred=hb["pred"])
        return hb["pred"].cpu().numpy()

    def f_predict_proba(self, X):
        logits = self._forward(X, train=False)
        n, C = logits.shape
        probs = self.l_out._mat(("probs", n), n, C)
        ops.softmax_ce(logits, probs=probs)
        return probs.cpu().numpy()

    # -------------------------------------------------------------------- fit
    def _batch(self, X, idx):
        if isinstance(X, (CSRMatrix, RowBlockedCSR)):
            return X.gather_rows_device(idx)
        return ops.gather_rows(X, torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int32)).to(self.device))

    def prepare(self, X_train, Y_train):
        if self.device.type != "cuda":
            raise RuntimeError("MLP runs on a CUDA device only; there is no CPU fallback")
        in_size = X_train.shape[1]
        Y_train = np.asarray(Y_train)
        out_size = len(set(Y_train.tolist()))                               # mlp.py:135
        if Y_train.min() < 0 or Y_train.max() >= out_size:
            raise ValueError("labels must be 0..%d (mlp.py:135 sizes the output by the number of distinct labels)"
                             % (out_size - 1))
        logger.info('output size is %d', out_size)
        if not self.hidden_layer_size:
            self.hidden_layer_size = min(5 * out_size, int(in_size / 20))  # mlp.py:141
        logger.info('input layer size: %d, hidden layer size: %d, output layer size: %d',
                    in_size, self.hidden_layer_size, out_size)
        if not is_sparse(X_train):
            logger.info('input matrix is not sparse!')                      # mlp.py:145
        self._build(is_sparse(X_train), in_size, out_size)
        self.Xd_train = self._to_device(X_train)                            # mlp.py:253 astype('float32')
        self.y_train_host = np.ascontiguousarray(Y_train, dtype=np.int32)
        self.y_train_dev = self._labels(Y_train)
        self._shuffle_rng = np.random.RandomState(self.seed) if self.seed is not None else np.random
        return self

    def train_epoch(self):
        """One pass of iterate_minibatches(X_train, Y_train, batch_size, shuffle=True) (mlp.py:268-270)."""
        n = self.Xd_train.shape[0]
        indices = np.arange(n)
        self._shuffle_rng.shuffle(indices)                                  # mlp.py:84-85
        n_batches = 0
        for start in range(0, n - self.batch_size + 1, self.batch_size):    # mlp.py:86: full batches only
            excerpt = indices[start:start + self.batch_size]
            x_batch = self._batch(self.Xd_train, excerpt)
            y_batch = self._labels(self.y_train_host[excerpt])
            self.f_train(x_batch, y_batch)
            n_batches += 1
        return n_batches

    def fit(self, X_train, Y_train, X_dev, Y_dev):
        logger.info('building the network... hidden:%s', self.add_hidden)
        self.prepare(X_train, Y_train)
        Xd_dev = self._to_device(X_dev)
        y_dev = self._labels(Y_dev)
        logger.info('training (n_epochs, batch_size) = (%s, %s)', self.n_epochs, self.batch_size)
        best_params = None
        best_val_acc = 0.0
        n_validation_down = 0
        self.history = []
        for n in range(self.n_epochs):                                      # mlp.py:267
            n_batches = self.train_epoch()
            if n_batches == 0:
                raise ValueError("batch_size %d exceeds the %d training rows: no full minibatch (mlp.py:86)"
                                 % (self.batch_size, self.Xd_train.shape[0]))
            l_train, acc_train = self.train_results()
            l_val, acc_val = self.f_val(Xd_dev, y_dev)                      # mlp.py:271 (the last call is what counts)
            self.history.append((l_train, acc_train, l_val, acc_val))
            if acc_val > best_val_acc:                                      # mlp.py:273-277
                best_val_acc = acc_val
                best_params = L.get_all_param_values(self.l_out)
                n_validation_down = 0
            else:
                n_validation_down += 1                                      # mlp.py:280
            logger.info('epoch %d ,train_loss %s ,acc %s ,val_loss %s ,acc %s,best_val_acc %s',
                        n, l_train, acc_train, l_val, acc_val, best_val_acc)
            if n_validation_down > self.early_stopping_max_down:            # mlp.py:282-284
                logger.info('validation results went down. early stopping ...')
                break
        if best_params is not None:
            L.set_all_param_values(self.l_out, best_params)                 # mlp.py:286
        logger.info('***************** final results based on best validation **************')
        l_val, acc_val = self.f_val(Xd_dev, y_dev)                          # mlp.py:289
        logger.info('Best dev acc: %f', acc_val)
        self.best_val = (l_val, acc_val)
        return self

    # ---------------------------------------------------------------- predict
    def predict(self, X_test):                                              # mlp.py:292-294
        return self.f_predict(self._to_device(X_test))

    def predict_proba(self, X_test):                                        # mlp.py:296-298
        return self.f_predict_proba(self._to_device(X_test))

    def accuracy(self, X_test, Y_test):                                     # mlp.py:300-307
        _loss, acc = self.f_val(self._to_device(X_test), self._labels(Y_test))
        return acc

    def score(self, X_test, Y_test):                                        # mlp.py:309-310
        return self.accuracy(X_test, Y_test)

    def get_embedding(self, X):                                             # mlp.py:311-312, :199-200
        if not self.add_hidden:
            raise ValueError("get_embedding needs add_hidden=True (mlp.py:198-200)")
        x = self._to_device(X)
        for ly in self.layers:
            x = ly.get_output_for(x, deterministic=True)
            if ly is self.l_hid1:
                return x.cpu().numpy()

    def get_param_values(self):
        return L.get_all_param_values(self.l_out)

    def get_grad_values(self):
        return [g.detach().cpu().numpy().copy() for g in self.grads]


class _DenseDropout(L.Layer):
    """lasagne.layers.dropout on a dense input (mlp.py:167-168); mask from torch's generator."""

    def __init__(self, incoming, p=0.5, **kwargs):
        super().__init__(incoming, **kwargs)
        self.p = float(p)

    def get_output_for(self, input, deterministic=False, **kwargs):
        if deterministic or self.p == 0:
            return input
        retain = 1.0 - self.p
        out = ops.alloc_mat(input.shape[0], input.shape[1], input.device)
        torch.mul(input, (torch.rand(input.shape, device=input.device) < retain).to(torch.float32) / retain, out=out)
        return out
