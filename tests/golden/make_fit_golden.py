"""Generates tests/golden/fit_golden.npz by running the REFERENCE'S OWN ``MLPCONV.fit``
(/root/reference/mlpconv.py:152-318, imported as a module, never copied) end to end: network construction
(:196-217), loss / elastic-net composition (:228-245), accuracy (:252), ``lasagne.updates.adam`` on the
trainable parameters (:263), the compiled functions (:265-268), the epoch loop with validation every 10
epochs, best-parameter snapshot, pickle and restore (:288-317), then ``predict`` / ``predict_proba`` (:320-336).

Theano and Lasagne are not installable here, so the module runs on a LAZY STUB of the two packages defined in
this file: symbolic variables are expression nodes, ``theano.function`` evaluates them on torch-CPU float32
tensors with torch autograd standing in for ``theano.grad``.  What the stub stands for is third-party behaviour
([3P] in SURVEY.md) and is written from the packages' documented semantics:
  S.dot(sparse, dense)            scipy ``csr @ dense`` forward (sparsetools csr_matvecs), ``x.T @ g`` backward
  T.dot                           BLAS sgemm (torch.mm)
  rectify / softmax               0.5*(x+|x|);  exp(x - rowmax) / rowsum
  categorical_crossentropy        -log(p[i, y_i]) for integer targets
  regularize_layer_params(l, p)   sum of p(W) over the REGULARIZABLE parameters of that one layer (W, not b)
  l1 / l2                         sum|x| / sum x^2
  GlorotUniform                   U(-a, a), a = sqrt(6 / (fan_in + fan_out)); biases 0
  lasagne.updates.adam            t+=1; a_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v moments; p -= a_t*m/(sqrt(v)+eps)
Everything ELSE that happens -- which layers exist, what feeds what, which coefficient multiplies which
penalty, that eval_loss carries the penalties too, what f_train returns, when validation runs, which parameters
are restored -- is the reference's code executing.

Run in the build container only:  python tests/golden/make_fit_golden.py
"""
import builtins
import collections
import importlib
import os
import sys
import tempfile
import types

import numpy as np
import scipy.sparse as sp
import torch

REF_DIR = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fit_golden.npz")
F32 = torch.float32


# ----------------------------------------------------------------------------- lazy expression nodes
def _const(v):
    return v if isinstance(v, Node) else Node(lambda env, v=v: torch.tensor(v, dtype=F32) if not torch.is_tensor(v) else v)


class Node(object):
    def __init__(self, fn, *parents):
        self.fn, self.parents = fn, parents

    def ev(self, env):
        k = id(self)
        if k not in env:
            env[k] = self.fn(env, *[p.ev(env) for p in self.parents])
        return env[k]

    def _bin(self, other, op, rev=False):
        o = _const(other)
        a, b = (o, self) if rev else (self, o)
        return Node(lambda env, x, y: op(x, y), a, b)

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, torch.sub, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.div)
    def __rtruediv__(self, o): return self._bin(o, torch.div, True)
    __div__, __rdiv__ = __truediv__, __rtruediv__
    def __pow__(self, o): return self._bin(o, torch.pow)
    def __rpow__(self, o): return self._bin(o, torch.pow, True)
    def __neg__(self): return Node(lambda env, x: -x, self)

    def __getitem__(self, key):
        if isinstance(key, tuple) and len(key) == 2 and isinstance(key[0], Node) and key[1] == slice(None):
            return Node(lambda env, x, i: x.index_select(0, i.long()), self, key[0])       # activation[target_indices, :]
        raise NotImplementedError("stub indexing %r" % (key,))

    def mean(self): return Node(lambda env, x: x.float().mean(), self)
    def sum(self): return Node(lambda env, x: x.sum(), self)
    def argmax(self, axis): return Node(lambda env, x: x.argmax(axis), self)

    def dimshuffle(self, *pattern):
        assert pattern == ("x", 0)
        return Node(lambda env, x: x[None, :], self)


class Input(Node):
    def __init__(self, name=None):
        Node.__init__(self, None)
        self.name = name

    def ev(self, env):
        return env[("in", id(self))]


class SparseVariable(Input):
    pass


class Shared(Node):
    def __init__(self, value, name=None):
        Node.__init__(self, None)
        self.value = torch.as_tensor(np.asarray(value, dtype=np.float32)).clone()
        self.name = name

    def ev(self, env):
        return env.get(("sh", id(self)), self.value)

    def get_value(self): return self.value.detach().numpy().copy()
    def set_value(self, v): self.value = torch.as_tensor(np.asarray(v, dtype=np.float32)).clone()


class _SparseDot(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dense, holder):
        ctx.m = holder[0]
        return torch.from_numpy(np.asarray(ctx.m @ dense.detach().numpy(), dtype=np.float32))

    @staticmethod
    def backward(ctx, g):
        return torch.from_numpy(np.asarray(sp.csc_matrix(ctx.m).T @ g.numpy(), dtype=np.float32)), None


def s_dot(a, b):
    b = _const(b)
    if isinstance(a, Node):
        return Node(lambda env, m, d: _SparseDot.apply(d, [m]), a, b)
    mat = sp.csr_matrix(a)                                  # a real scipy matrix (the layers hold H itself)
    return Node(lambda env, d: _SparseDot.apply(d, [mat]), b)


CALLS = collections.defaultdict(list)      # function name -> list of outputs per call


def install_stubs(rng):
    theano = types.ModuleType("theano")
    tensor = types.ModuleType("theano.tensor")
    sparse = types.ModuleType("theano.sparse")
    tensor.dot = lambda a, b: Node(lambda env, x, y: torch.mm(x, y), _const(a), _const(b))
    tensor.ivector = lambda *a, **k: Input()
    tensor.matrix = lambda *a, **k: Input()
    tensor.constant = lambda v, name=None: _const(float(v))
    tensor.sqrt = lambda x: Node(lambda env, v: torch.sqrt(v), _const(x))
    tensor.mean = lambda x: _const(x).mean()
    tensor.eq = lambda a, b: Node(lambda env, x, y: (x.long() == y.long()), _const(a), _const(b))
    sparse.SparseVariable = SparseVariable
    sparse.SparseConstant = SparseVariable
    sparse.sharedvar = types.SimpleNamespace(SparseTensorSharedVariable=SparseVariable)
    sparse.csr_matrix = lambda name=None, dtype=None: SparseVariable(name)
    sparse.dot = s_dot
    theano.tensor, theano.sparse = tensor, sparse
    theano.shared = lambda v, **k: Shared(v)
    n_fn = [0]

    def function(inputs, outputs, updates=None, on_unused_input=None, **kw):
        idx = n_fn[0]
        n_fn[0] += 1
        outs = outputs if isinstance(outputs, (list, tuple)) else [outputs]

        def call(*args):
            env = {}
            for sym, val in zip(inputs, args):
                env[("in", id(sym))] = val if sp.issparse(val) else torch.as_tensor(np.asarray(val))
            vals = [o.ev(env) for o in outs]
            new = [(s, e.ev(env)) for s, e in (updates or {}).items()]
            for s, v in new:
                s.value = v.detach().clone()
            res = [v.detach().numpy().copy() for v in vals]
            CALLS[idx].append([np.asarray(r) for r in res])
            return res if isinstance(outputs, (list, tuple)) else res[0]
        return call
    theano.function = function

    def grad(loss, params):
        cache = {}

        def make(i):
            def fn(env):
                key = ("grads", id(loss))
                if key not in env:
                    leaves = []
                    for p in params:
                        leaf = p.value.detach().clone().requires_grad_(True)
                        leaves.append(leaf)
                    env2 = {k: v for k, v in env.items() if isinstance(k, tuple) and k[0] == "in"}
                    for p, leaf in zip(params, leaves):
                        env2[("sh", id(p))] = leaf
                    env[key] = torch.autograd.grad(loss.ev(env2), leaves)
                return env[key][i]
            return Node(fn)
        return [make(i) for i in range(len(params))]
    theano.grad = grad

    lasagne = types.ModuleType("lasagne")
    layers = types.ModuleType("lasagne.layers")
    reg = types.ModuleType("lasagne.regularization")

    class Layer(object):
        def __init__(self, incoming, name=None):
            self.input_layer = incoming if isinstance(incoming, Layer) else None
            self.input_shape = incoming.output_shape if isinstance(incoming, Layer) else tuple(incoming)
            self.params = collections.OrderedDict()

        @property
        def output_shape(self):
            return self.input_shape

    class InputLayer(Layer):
        def __init__(self, shape, input_var=None, **k):
            Layer.__init__(self, shape)
            self.input_var = input_var

    class DenseLayer(Layer):
        def __init__(self, incoming, num_units, W=None, b=None, nonlinearity=None, **k):
            Layer.__init__(self, incoming)
            self.num_units = num_units
            fan_in = self.input_shape[1]
            a = np.sqrt(6.0 / (fan_in + num_units))
            self.W = Shared(rng.uniform(-a, a, size=(fan_in, num_units)).astype(np.float32), "W")
            self.b = Shared(np.zeros(num_units, np.float32), "b")
            self.params[self.W] = {"trainable", "regularizable"}
            self.params[self.b] = {"trainable"}
            self.nonlinearity = nonlinearity if nonlinearity is not None else (lambda x: x)

        @property
        def output_shape(self):
            return (self.input_shape[0], self.num_units)

    class DropoutLayer(Layer):
        pass

    def all_layers(layer):
        chain = []
        while layer is not None:
            chain.append(layer)
            layer = layer.input_layer
        return chain[::-1]

    def get_output(layer, inputs=None, **kwargs):
        x = inputs
        for ly in all_layers(layer):
            if isinstance(ly, InputLayer):
                x = ly.input_var if x is None else x
            else:
                x = ly.get_output_for(x, **kwargs)
        return x

    def get_all_params(layer, **tags):
        out = []
        for ly in all_layers(layer):
            for p, t in ly.params.items():
                if all((k in t) == v for k, v in tags.items()):
                    out.append(p)
        return out
    layers.Layer, layers.InputLayer, layers.DenseLayer, layers.DropoutLayer = Layer, InputLayer, DenseLayer, DropoutLayer
    layers.get_output, layers.get_all_params = get_output, get_all_params
    layers.get_all_param_values = lambda layer: [p.get_value() for p in get_all_params(layer)]

    def set_all_param_values(layer, values):
        for p, v in zip(get_all_params(layer), values):
            p.set_value(v)
    layers.set_all_param_values = set_all_param_values
    layers.dropout = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("dropout is off in this run"))

    reg.l1 = lambda x: Node(lambda env, v: v.abs().sum(), x)
    reg.l2 = lambda x: Node(lambda env, v: (v * v).sum(), x)

    def regularize_layer_params(layer, penalty, tags={"regularizable": True}, **k):
        tot = None
        for p, t in layer.params.items():
            if "regularizable" in t:
                tot = penalty(p) if tot is None else tot + penalty(p)
        return tot
    reg.regularize_layer_params = regularize_layer_params
    reg.regularize_layer_params_weighted = None

    nl = types.SimpleNamespace(
        rectify=lambda x: Node(lambda env, v: 0.5 * (v + v.abs()), x),
        sigmoid=lambda x: Node(lambda env, v: torch.sigmoid(v), x),
        tanh=lambda x: Node(lambda env, v: torch.tanh(v), x),
        softmax=lambda x: Node(lambda env, v: (lambda e: e / e.sum(1, keepdim=True))(torch.exp(v - v.max(1, keepdim=True).values)), x))
    objectives = types.SimpleNamespace(
        categorical_crossentropy=lambda p, t: Node(lambda env, pv, tv: -torch.log(pv[torch.arange(pv.shape[0]), tv.long()]), p, t))

    def adam(loss, params, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
        grads = grad(loss, params)
        t_prev = Shared(np.float32(0.0))
        updates = collections.OrderedDict()
        one = _const(1.0)
        t = t_prev + 1
        a_t = learning_rate * tensor.sqrt(one - beta2 ** t) / (one - beta1 ** t)
        for p, g_t in zip(params, grads):
            m_prev, v_prev = Shared(np.zeros_like(p.get_value())), Shared(np.zeros_like(p.get_value()))
            m_t = beta1 * m_prev + (one - beta1) * g_t
            v_t = beta2 * v_prev + (one - beta2) * g_t ** 2
            step = a_t * m_t / (tensor.sqrt(v_t) + epsilon)
            updates[m_prev], updates[v_prev], updates[p] = m_t, v_t, p - step
        updates[t_prev] = t
        return updates
    lasagne.layers, lasagne.regularization, lasagne.nonlinearities, lasagne.objectives = layers, reg, nl, objectives
    lasagne.updates = types.SimpleNamespace(adam=adam)
    lasagne.init = types.SimpleNamespace(GlorotUniform=lambda *a, **k: None)
    sys.modules.update({"theano": theano, "theano.tensor": tensor, "theano.sparse": sparse, "lasagne": lasagne,
                        "lasagne.layers": layers, "lasagne.regularization": reg})
    return layers


def main():
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from make_layers_golden import mention_graph, random_csr, reference_ahat       # the reference's own A_hat
    for m in [k for k in sys.modules if k.split(".")[0] in ("theano", "lasagne", "lasagne_layers")]:
        del sys.modules[m]
    rng = np.random.RandomState(123)
    layers = install_stubs(np.random.RandomState(7))
    builtins.xrange = range
    sys.maxint = sys.maxsize
    sys.path.insert(0, REF_DIR)
    try:
        ref = importlib.import_module("mlpconv")
    finally:
        sys.path.remove(REF_DIR)
    assert os.path.realpath(ref.__file__).startswith(REF_DIR)

    n, V, hid, C = 260, 150, 24, 9
    g = mention_graph(rng, n, 7, isolated=(4,))
    _, H = reference_ahat(g, n)
    X = random_csr(rng, n, V, 10, empty_rows=(11,))
    Y = rng.randint(0, C, size=n)
    Y[:C] = np.arange(C)
    n_train, n_dev = 160, 50
    train_idx = rng.choice(n_train, size=n_train, replace=True).astype(np.int32)     # tensormain.py:226: with replacement
    dev_idx = np.arange(n_train, n_train + n_dev, dtype=np.int32)
    test_idx = np.arange(n_train + n_dev, n, dtype=np.int32)
    regul = [3e-4, 7e-4]                       # (out, hid), deliberately different: mlpconv.py:237 unpack order
    clf = ref.MLPCONV(n_epochs=25, batch_size=10, init_parameters=None, complete_prob=False, add_hidden=True,
                      regul_coefs=regul, save_results=False, hidden_layer_size=hid, drop_out=False,
                      dropout_coefs=[0.5, 0.5], early_stopping_max_down=1, loss_name='log', nonlinearity='rectify',
                      dtype='float32')
    created = []
    orig_dense_init = layers.DenseLayer.__init__

    def recording_init(self, *a, **k):
        orig_dense_init(self, *a, **k)
        created.append((self.W.get_value(), self.b.get_value()))
    layers.DenseLayer.__init__ = recording_init
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "data"))
        os.chdir(tmp)
        try:
            clf.fit(X, train_idx, dev_idx, test_idx, Y, H)                          # mlpconv.py:152-318
            pkl = [f for f in os.listdir("data") if f.endswith(".pkl")]
            import pickle
            best = pickle.load(open(os.path.join("data", pkl[0]), "rb"))
        finally:
            os.chdir(cwd)
    init = [a for wb in created for a in wb]
    out = {
        "X__indptr": X.indptr.astype(np.int32), "X__indices": X.indices.astype(np.int32), "X__data": X.data,
        "X__shape": np.array(X.shape, np.int64), "H__indptr": H.indptr.astype(np.int32), "H__indices": H.indices.astype(np.int32),
        "H__data": H.data.astype(np.float32), "Y": Y.astype(np.int64), "train_idx": train_idx, "dev_idx": dev_idx,
        "test_idx": test_idx, "regul_coefs": np.array(regul), "hidden": np.array([hid]), "pickle_name": np.array(pkl[0]),
        "f_train": np.array([[float(c[0]), float(c[1])] for c in CALLS[0]]),        # [loss, acc] per epoch (:295)
        "f_val": np.array([[float(c[0]), float(c[1])] for c in CALLS[1]]),          # epochs 0, 10, 20 + final (:297, :317)
        "predict_test": np.asarray(clf.predict("test")).astype(np.int64),           # :320-327
        "proba_dev": np.asarray(clf.predict_proba("dev")).astype(np.float32),       # :329-336
        "accuracy_test": np.array([float(clf.accuracy("test", Y[test_idx].astype("int32")))]),
    }
    for i, a in enumerate(init):
        out["init_%d" % i] = a
    for i, a in enumerate(best):
        out["best_%d" % i] = np.asarray(a, np.float32)
    for i, a in enumerate(layers.get_all_param_values(clf.l_out)):
        out["final_%d" % i] = a
    print("f_train first/last", out["f_train"][0], out["f_train"][-1], "f_val", out["f_val"])
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
