"""``MLPCONV`` -- the full-batch GCN of the reference, on B200.

Same constructor keywords and the same ``fit / predict / predict_proba /
accuracy`` signatures as /root/reference/mlpconv.py:121-346, so a
``tensormain.main_mlpconv``-style driver (tensormain.py:232-244) runs unchanged.
One epoch = one ``f_train`` call = full-graph forward + backward + Adam step
(mlpconv.py:293-295), enqueued on one CUDA stream through libgcg.so and replayed
as a CUDA graph.

Extensions (BASELINE.json north_star): ``n_layers`` (>2 adds hidden conv layers)
and ``highway`` (gates those hidden layers).  ``n_layers=2, highway=False`` is
exactly the reference network (mlpconv.py:196-217).
"""
from __future__ import annotations

import logging
import os
import pickle
import sys

import numpy as np
import torch

from . import lasagne_layers as L
from . import ops
from .sparse import as_csr

logger = logging.getLogger("graphconvgeo_b200")


class MLPCONV:
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
                 dropout_coefs=[0.5, 0.5],
                 early_stopping_max_down=100000,
                 loss_name='log',
                 nonlinearity='rectify',
                 dtype='float32',
                 # ---- extensions ----
                 n_layers=2,
                 highway=False,
                 device='cuda',
                 seed=None,
                 cuda_graph=True,
                 model_dir=None,
                 learning_rate=4e-3,
                 reorder="auto",
                 native_epoch=None):
        # mlpconv.py:136-150
        self.n_epochs = n_epochs
        self.batch_size = batch_size          # accepted and ignored: full batch (mlpconv.py:294)
        self.init_parameters = init_parameters
        self.complete_prob = complete_prob
        self.add_hidden = add_hidden
        self.regul_coefs = regul_coefs
        self.save_results = save_results
        self.hidden_layer_size = hidden_layer_size
        self.drop_out = drop_out
        self.dropout_coefs = dropout_coefs
        self.early_stopping_max_down = early_stopping_max_down
        self.loss_name = loss_name
        # the reference hard-codes 'rectify' here (mlpconv.py:149); we honour the argument
        self.nonlinearity = nonlinearity if nonlinearity in ('rectify', 'sigmoid', 'tanh') else 'rectify'
        self.dtype = dtype
        self.n_layers = int(n_layers)
        self.highway = bool(highway)
        self.device = torch.device(device)
        self.seed = seed
        self.cuda_graph = cuda_graph
        # the epoch as a native gcg_epoch object (see f_train); None = the GCG_NATIVE_EPOCH environment switch
        self.native_epoch = (os.environ.get("GCG_NATIVE_EPOCH", "1") != "0") if native_epoch is None else bool(native_epoch)
        self.model_dir = model_dir
        self.learning_rate = learning_rate
        self.reorder = reorder      # None | "auto" | "labels" | "degree" | explicit permutation (new -> old)
        if complete_prob:
            raise NotImplementedError("complete_prob=True (dense label distributions) is unused by the "
                                      "reference driver (tensormain.py:232) and not on the hot path")
        if dtype != 'float32':
            raise ValueError("the hot path is float32 (mlpconv.py:136, tensormain.py:212)")
        if loss_name != 'log':
            raise ValueError("only loss_name='log' exists in the reference (mlpconv.py:228)")
        assert self.n_layers >= 2

    # ------------------------------------------------------------------ build
    def _build(self, X, H, in_size, out_size):
        rng = np.random.RandomState(self.seed) if self.seed is not None else None
        drop_out_hid, drop_out_in = self.dropout_coefs                     # mlpconv.py:155
        l_in = L.InputLayer(shape=(None, in_size), input_var=X, device=self.device)   # :196-197
        if self.drop_out:
            l_in = L.SparseInputDropoutLayer(l_in, p=drop_out_in)          # :199-201
        l_hid = L.SparseConvolutionDenseLayer(l_in, H=H, num_units=self.hidden_layer_size,
                                              nonlinearity=self.nonlinearity,
                                              W=L.GlorotUniform(), rng=rng)  # :205-209
        self.l_hid1 = l_hid
        Hd = l_hid.H
        if self.drop_out:
            l_hid = DropoutLayer(l_hid, p=drop_out_hid)                    # :210-211
        for _ in range(self.n_layers - 2):                                 # extension
            cls = L.HighwayConvolutionDenseLayer if self.highway else L.ConvolutionDenseLayer
            l_hid = cls(l_hid, H=Hd, num_units=self.hidden_layer_size, nonlinearity=self.nonlinearity,
                        W=L.GlorotUniform(), rng=rng)
        self.l_out = L.ConvolutionDenseLayer(l_hid, H=Hd, num_units=out_size,
                                             nonlinearity=L.nonlinearities.softmax, rng=rng)   # :213-216
        self.layers = [ly for ly in L.get_all_layers(self.l_out) if not isinstance(ly, L.InputLayer)]
        if self.init_parameters is not None:
            L.set_all_param_values(self.l_out, self.init_parameters)
        # gradients + optimiser state (lasagne.updates.adam, :263) and regularisation (:235-245)
        regul_coef_out, regul_coef_hid = self.regul_coefs                  # :237
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
        self.params, self.grads, self.reg = params, grads, reg
        self.adam = ops.Adam(params, grads, reg, lr=self.learning_rate, beta1=0.9, beta2=0.999, eps=1e-8)
        self.elastic = ops.ElasticNet(params, reg)

    # -------------------------------------------------------------- functions
    def _forward(self, ti, train):
        x = self.Xd
        for ly in self.layers:
            if ly is self.l_out:
                x = ly.get_output_for(x, target_indices=ti, logits=True, train=train, deterministic=not train)
            else:
                x = ly.get_output_for(x, train=train, deterministic=not train)
        return x

    def _head_buffers(self, n, C):
        key = (n, C)
        hb = self._heads.get(key)
        if hb is None:
            d = self.device
            hb = dict(ce=torch.empty(n, dtype=torch.float32, device=d),
                      hit=torch.empty(n, dtype=torch.float32, device=d),
                      pred=torch.empty(n, dtype=torch.int64, device=d),
                      out=torch.zeros(2, dtype=torch.float32, device=d))
            self._heads[key] = hb
        return hb

    def _train_step_enqueue(self):
        """f_train (mlpconv.py:265, :295) enqueued on the current stream; results stay on the device."""
        ti, y = self.ti_train, self.y_train_dev
        logits = self._forward(ti, train=True)
        n, C = logits.shape
        hb = self._head_buffers(n, C)
        G = self.l_out._mat("G", n, C)
        ops.softmax_ce(logits, y=y, grad=G, ce=hb["ce"], hit=hb["hit"], denom=n)
        ops.sum_scaled(hb["ce"], 1.0 / n, out=hb["out"][0:1])
        ops.sum_scaled(hb["hit"], 1.0 / n, out=hb["out"][1:2])
        # backward: theano.grad through the chain (mlpconv.py:263)
        grad = G
        preact = False
        for i in range(len(self.layers) - 1, -1, -1):
            ly = self.layers[i]
            if isinstance(ly, DropoutLayer):
                grad = ly.backward(grad)
                preact = False
                continue
            prev = self.layers[i - 1] if i > 0 else None
            mask = None
            if prev is not None and type(prev) in (L.SparseConvolutionDenseLayer, L.ConvolutionDenseLayer) \
                    and prev.nonlinearity in ("rectify", "tanh"):
                mask = (prev._out, prev.nonlinearity)      # fuse prev's act' into this layer's dH product
            if isinstance(ly, L.SparseConvolutionDenseLayer):
                ly.backward(grad, preact=preact)
                break
            grad = ly.backward(grad, preact=preact, input_mask=mask) if not isinstance(ly, L.HighwayConvolutionDenseLayer) \
                else ly.backward(grad, input_mask=mask)
            preact = mask is not None
        self.adam.step()       # also leaves the elastic-net penalty of the pre-update params in adam.reg_out
        self._train_hb = hb

    def f_train(self):
        """One epoch (mlpconv.py:295).  The first call runs the layer code eagerly (buffers get allocated); after
        that the epoch is a compiled object, as the reference's Theano function is: with ``native_epoch`` the
        second call records the step into a gcg_epoch (C++ list of every libgcg call, include/gcg.h) which later
        calls replay with one C call -- optionally itself captured in a CUDA graph; without it the layer code is
        captured in a CUDA graph directly."""
        if self._graph is not None:
            self._graph.replay()
        elif self._program is not None:
            self._program.run()
        elif self.native_epoch and self._steps_done >= 1 and not self.drop_out:
            prog = ops.EpochProgram()
            with prog.record():
                self._train_step_enqueue()
            self._program = prog
            self._steps_done += 1
            if self.cuda_graph:
                self._capture()
        else:
            self._train_step_enqueue()
            self._steps_done += 1
            if self.cuda_graph and not self.native_epoch and self._steps_done == 1 and not self.drop_out:
                self._capture()
        return self._train_hb

    def _capture(self):
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            if self._program is not None:
                self._program.run()
            else:
                self._train_step_enqueue()
        self._graph = g

    def train_results(self):
        """(loss, acc) of the last f_train as host scalars (synchronises)."""
        hb = self._train_hb
        ce_acc = hb["out"].cpu().numpy()
        reg = float(self.adam.reg_out.item())
        return float(np.float32(ce_acc[0]) + np.float32(reg)), float(ce_acc[1])

    def f_val(self, y, ti):
        """[eval_loss, eval_acc] (mlpconv.py:266): deterministic forward + CE + the same penalty."""
        logits = self._forward(ti, train=False)
        n, C = logits.shape
        hb = self._head_buffers(n, C)
        ops.softmax_ce(logits, y=y, ce=hb["ce"], hit=hb["hit"], denom=n)
        ops.sum_scaled(hb["ce"], 1.0 / n, out=hb["out"][0:1])
        ops.sum_scaled(hb["hit"], 1.0 / n, out=hb["out"][1:2])
        reg = self.elastic()
        o = hb["out"].cpu().numpy()
        return float(np.float32(o[0]) + np.float32(reg.item())), float(o[1])

    def f_predict(self, ti):
        logits = self._forward(ti, train=False)
        n, C = logits.shape
        hb = self._head_buffers(n, C)
        ops.softmax_ce(logits, pred=hb["pred"])
        return hb["pred"].cpu().numpy()

    def f_predict_proba(self, ti):
        logits = self._forward(ti, train=False)
        n, C = logits.shape
        probs = self.l_out._mat(("probs", n), n, C)
        ops.softmax_ce(logits, probs=probs)
        return probs.cpu().numpy()

    # -------------------------------------------------------------------- fit
    def prepare(self, X, train_indices, dev_indices, test_indices, Y, H):
        """Everything fit() does before the epoch loop (mlpconv.py:152-287)."""
        if not self.device.type == "cuda":
            raise RuntimeError("MLPCONV runs on a CUDA device only; there is no CPU fallback")
        logger.info('building the network... hidden:%s', self.add_hidden)
        in_size = X.shape[1]
        Y = np.asarray(Y)
        Y_train = Y[train_indices]                                         # mlpconv.py:162-164
        Y_dev = Y[dev_indices]
        out_size = int(np.max(Y)) + 1                                      # :165
        logger.info('output size is %d', out_size)
        self.X = X
        self.train_indices = train_indices
        self.dev_indices = dev_indices
        self.test_indices = test_indices
        self.H = H
        logger.info('input layer size: %d, hidden layer size: %d, output layer size: %d, dropout %s, regul %s, dtype %s',
                    in_size, self.hidden_layer_size, out_size, str(self.dropout_coefs), str(self.regul_coefs), self.dtype)
        self.Xd = as_csr(X, self.device, long_row_threshold=1024)
        Hd = as_csr(H, self.device)
        # node reordering (performance only; the net is permutation equivariant): nodes of the same
        # region become neighbours in memory, so gathered rows of A_hat.H share L2 lines
        n_nodes = Hd.shape[0]
        mode = self.reorder
        if isinstance(mode, str) and mode == "auto":
            mode = "labels" if n_nodes >= 100000 else None
        self.node_order = None                       # new position -> original node id
        node_map = lambda idx: np.asarray(idx)
        if mode is not None:
            if isinstance(mode, str) and mode == "labels":
                order = np.argsort(Y[:n_nodes], kind="stable").astype(np.int32)
            elif isinstance(mode, str) and mode == "degree":
                deg = np.diff(Hd._host_arrays()[0])
                order = np.argsort(-deg, kind="stable").astype(np.int32)
            else:
                order = np.ascontiguousarray(np.asarray(mode), dtype=np.int32)
            inv = np.empty(n_nodes, np.int32)
            inv[order] = np.arange(n_nodes, dtype=np.int32)
            self.Xd = self.Xd.permute(order)
            Hd = Hd.permute(order, col_map=inv)
            self.node_order, self.node_inverse = order, inv
            node_map = lambda idx: inv[np.asarray(idx)]
        self._node_map = node_map
        self._build(self.Xd, Hd, in_size, out_size)
        Hd = self.l_hid1.H
        self.ti = {"train": L.TargetIndices(node_map(train_indices), Hd),
                   "dev": L.TargetIndices(node_map(dev_indices), Hd),
                   "test": L.TargetIndices(node_map(test_indices), Hd)}
        self.ti_train = self.ti["train"]
        to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(self.device)
        self.y_train_dev = to_dev(Y_train)                                 # :276-277 int32
        self.y_dev_dev = to_dev(Y_dev)
        self._heads = {}
        self._graph = None
        self._program = None
        self._steps_done = 0
        self._train_hb = None
        return self

    def fit(self, X, train_indices, dev_indices, test_indices, Y, H):
        self.prepare(X, train_indices, dev_indices, test_indices, Y, H)
        logger.info('training (n_epochs, batch_size) = (%s, %s)', self.n_epochs, self.batch_size)
        best_params = None
        best_val_loss = sys.maxsize                                        # :289
        best_val_acc = 0.0
        n_validation_down = 0
        report_k_epoch = 10
        for n in range(self.n_epochs):                                     # :293
            self.f_train()                                                 # :295
            if n % report_k_epoch == 0:
                l_train, acc_train = self.train_results()
                if not np.isfinite(l_train):
                    # the reference has a NaN detector only in its MDN scripts (lang2loc_mdnshared.py:187-199);
                    # here a non-finite loss stops the fit instead of training on garbage for n_epochs
                    raise FloatingPointError("training loss is %r at epoch %d" % (l_train, n))
                l_val, acc_val = self.f_val(self.y_dev_dev, self.ti["dev"])   # :297
                if l_val < best_val_loss:
                    best_val_loss = l_val
                    best_val_acc = acc_val
                    best_params = L.get_all_param_values(self.l_out)       # :301
                    n_validation_down = 0
                else:
                    n_validation_down += 1                                 # :305
                logger.info('epoch %d ,train_loss %s ,acc %s ,val_loss %s ,acc %s,best_val_acc %s',
                            n, l_train, acc_train, l_val, acc_val, best_val_acc)
                if n_validation_down > self.early_stopping_max_down:       # :307
                    logger.info('validation results went down. early stopping ...')
                    break
        if best_params is not None:
            if self.model_dir is not None:                                 # :310-313
                os.makedirs(self.model_dir, exist_ok=True)
                model_file = os.path.join(self.model_dir, 'Xshape1_' + str(X.shape[1]) + '_hidden_' +
                                          str(self.hidden_layer_size) + '_regul_' + str(self.regul_coefs[0]) +
                                          '_drop_' + str(self.dropout_coefs[0]) + '.pkl')
                logger.info('storing best parameters in %s ...', model_file)
                with open(model_file, 'wb') as fout:
                    pickle.dump(best_params, fout)
            L.set_all_param_values(self.l_out, best_params)                # :314
        logger.info('***************** final results based on best validation **************')
        l_val, acc_val = self.f_val(self.y_dev_dev, self.ti["dev"])        # :317
        logger.info('Best dev acc: %f', acc_val)
        self.best_val = (l_val, acc_val)
        return self

    # ---------------------------------------------------------------- predict
    def _partition(self, dataset_partition):
        if dataset_partition not in ("train", "dev", "test"):
            raise ValueError("dataset_partition must be 'train', 'dev' or 'test'")
        return self.ti[dataset_partition]

    def predict(self, dataset_partition):                                  # mlpconv.py:320-327
        return self.f_predict(self._partition(dataset_partition))

    def predict_proba(self, dataset_partition):                            # :329-336
        return self.f_predict_proba(self._partition(dataset_partition))

    def accuracy(self, dataset_partition, y_true):                         # :338-346
        ti = self._partition(dataset_partition)
        y = torch.from_numpy(np.ascontiguousarray(y_true, dtype=np.int32)).to(self.device)
        _loss, _acc = self.f_val(y, ti)
        return _acc

    def score(self, dataset_partition, y_true):
        # the reference's score() passes a stray positional argument (mlpconv.py:348-349) and raises
        return self.accuracy(dataset_partition, y_true)

    def get_embedding(self, indices):
        """Hidden representation of ``indices`` (a stub in the reference, mlpconv.py:350-352)."""
        self._forward(self.ti["dev"], train=False)
        idx = torch.from_numpy(np.ascontiguousarray(self._node_map(indices), dtype=np.int32)).to(self.device)
        return ops.gather_rows(self.layers[-2]._out, idx).cpu().numpy()

    def node_rows(self, t):
        """A per-node device matrix (e.g. a layer's ``_out``) as a host array in ORIGINAL node order."""
        a = t.detach().cpu().numpy()
        return a if self.node_order is None else a[self.node_inverse]

    def host_inputs(self):
        """(X, A_hat) as scipy CSR matrices in the MODEL'S node order (after the locality reordering) -- what a
        checker needs to recompute rows of any layer on the host."""
        return self.Xd.to_scipy(), self.l_hid1.H.to_scipy()

    def get_param_values(self):
        return L.get_all_param_values(self.l_out)

    def get_grad_values(self):
        return [g.detach().cpu().numpy().copy() for g in self.grads]


class DropoutLayer(L.Layer):
    """lasagne.layers.dropout on a dense activation (mlpconv.py:210-211).  Off in every
    parity / bench configuration (the reference driver runs drop_out=False,
    tensormain.py:234); the mask comes from torch's generator, not Theano's."""

    def __init__(self, incoming, p=0.5, rescale=True, **kwargs):
        super().__init__(incoming, **kwargs)
        self.p, self.rescale = float(p), rescale
        self._mask = None
        self.nonlinearity = "identity"

    def get_output_for(self, input, deterministic=False, **kwargs):
        self._out = input
        if deterministic or self.p == 0:
            self._mask = None
            return input
        retain = 1.0 - self.p
        self._mask = (torch.rand(input.shape, device=input.device) < retain).to(torch.float32)
        if self.rescale:
            self._mask /= retain
        out = self._mat("out", *input.shape)
        torch.mul(input, self._mask, out=out)
        return out

    def backward(self, grad_output, **kwargs):
        if self._mask is None:
            return grad_output
        return grad_output.mul_(self._mask)
