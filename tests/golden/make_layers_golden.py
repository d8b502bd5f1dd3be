"""Generates tests/golden/layers_golden.npz by EXECUTING the reference's own source:

  * /root/reference/lasagne_layers.py is imported as a module (never copied) with stub ``theano`` /
    ``lasagne`` packages in ``sys.modules``; the three hot-path bodies
    ``SparseInputDenseLayer / SparseConvolutionDenseLayer / ConvolutionDenseLayer.get_output_for``
    (lasagne_layers.py:20-29, 53-71, 73-89) then run eagerly on NumPy / scipy inputs;
  * the A_hat statements of ``preprocess_data`` (tensormain.py:170-180) and the float32 cast of
    ``main_mlpconv`` (tensormain.py:221) are cut out with ``ast`` and executed on seeded networkx graphs;
  * ``geo_eval`` (tensormain.py:38-54) is cut out with ``ast`` and executed with a stub ``haversine``.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_layers_golden.py

What the stubs stand for (third-party behaviour, [3P] in SURVEY.md; everything else is reference code):
  S.dot(sparse, dense)   theano.sparse.basic.Dot.perform: ``x * y`` on the scipy matrix, i.e. sparsetools
                         ``csr_matvecs`` -> here ``x @ y`` on a scipy csr_matrix (the same routine)
  T.dot(a, b)            BLAS sgemm -> ``numpy.dot``
  b.dimshuffle('x', 0)   broadcastable row -> ``b[None, :]``
  nonlinearities         lasagne.nonlinearities.rectify = theano relu = 0.5*(x+|x|) (written that way here),
                         tanh, sigmoid = 1/(1+exp(-x)), softmax = exp(x - rowmax) / rowsum, identity
  S.SparseVariable ...   the isinstance targets of the sparse check -> scipy.sparse.spmatrix
  nx.adjacency_matrix    returned a scipy.sparse *matrix* in the reference's era (``*`` = matrix product,
                         ``.sum(axis=1)`` = numpy.matrix); networkx 3 returns a sparse *array*, so the result is
                         wrapped in csr_matrix before the reference statements run
  sp.errstate/sqrt/isinf SciPy's old re-exports of the NumPy functions (tensormain.py:175-177)
  haversine              great-circle km, R = 6371.0088 (the package's AVG_EARTH_RADIUS)
"""
import ast
import logging
import os
import sys
import types

import networkx as nx
import numpy as np
import scipy
import scipy.sparse

REF_DIR = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "layers_golden.npz")
F32 = np.float32


# ----------------------------------------------------------------------------- stub theano / lasagne
class _Param(np.ndarray):
    """A float32 array that answers ``dimshuffle('x', 0)`` like a Theano shared vector."""

    def dimshuffle(self, *pattern):
        assert pattern == ("x", 0)
        return np.asarray(self)[None, :]


def _param(a):
    return None if a is None else np.ascontiguousarray(a, dtype=F32).view(_Param)


def _rectify(x):
    return F32(0.5) * (x + np.abs(x))               # theano.tensor.nnet.relu(x, alpha=0)


def _sigmoid(x):
    return (F32(1) / (F32(1) + np.exp(-x, dtype=F32))).astype(F32)


def _softmax(x):
    e = np.exp(x - x.max(axis=1, keepdims=True), dtype=F32)
    return (e / e.sum(axis=1, keepdims=True, dtype=F32)).astype(F32)


NONLIN = {"rectify": _rectify, "tanh": lambda x: np.tanh(x, dtype=F32), "sigmoid": _sigmoid,
          "softmax": _softmax, "identity": lambda x: x}


def install_stubs():
    theano = types.ModuleType("theano")
    tensor = types.ModuleType("theano.tensor")
    sparse = types.ModuleType("theano.sparse")
    tensor.dot = lambda a, b: np.dot(a, b)
    tensor.constant = lambda v, name=None: F32(v)
    sparse.SparseVariable = scipy.sparse.spmatrix
    sparse.SparseConstant = scipy.sparse.spmatrix
    sparse.sharedvar = types.SimpleNamespace(SparseTensorSharedVariable=scipy.sparse.spmatrix)

    def s_dot(x, y):
        assert scipy.sparse.issparse(x)
        out = x @ y
        return out.toarray() if scipy.sparse.issparse(out) else np.asarray(out)
    sparse.dot = s_dot
    sparse.mul = lambda x, s: x.multiply(s)
    theano.tensor, theano.sparse = tensor, sparse

    lasagne = types.ModuleType("lasagne")
    reg = types.ModuleType("lasagne.regularization")
    for n in ("regularize_layer_params_weighted", "l2", "l1", "regularize_layer_params"):
        setattr(reg, n, None)
    layers = types.ModuleType("lasagne.layers")

    class Layer(object):
        def __init__(self, incoming, name=None):
            self.input_shape = getattr(incoming, "output_shape", incoming)

    class DenseLayer(Layer):
        def __init__(self, incoming, num_units, W=None, b=None, nonlinearity=None, **kwargs):
            Layer.__init__(self, incoming)
            self.num_units = num_units
            self.W, self.b = _param(W), _param(b)
            self.nonlinearity = NONLIN["identity"] if nonlinearity is None else nonlinearity

    class DropoutLayer(Layer):
        def __init__(self, incoming, p=0.5, rescale=True, **kwargs):
            Layer.__init__(self, incoming)
            self.p, self.rescale = p, rescale

    layers.Layer, layers.DenseLayer, layers.DropoutLayer = Layer, DenseLayer, DropoutLayer
    lasagne.regularization, lasagne.layers = reg, layers
    sys.modules.update({"theano": theano, "theano.tensor": tensor, "theano.sparse": sparse, "lasagne": lasagne,
                        "lasagne.regularization": reg, "lasagne.layers": layers})


def import_reference_layers():
    install_stubs()
    sys.path.insert(0, REF_DIR)
    try:
        import importlib
        mod = importlib.import_module("lasagne_layers")
    finally:
        sys.path.remove(REF_DIR)
    assert os.path.realpath(mod.__file__).startswith(REF_DIR)
    return mod


# ----------------------------------------------------------------------------- reference statements via ast
def _function_node(path, name):
    tree = ast.parse(open(path).read())
    return next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)


def reference_ahat(graph, n):
    """tensormain.py:170-180 executed verbatim (then the cast of :221)."""
    path = os.path.join(REF_DIR, "tensormain.py")
    fn = _function_node(path, "preprocess_data")
    stmts = [s for s in fn.body if 170 <= s.lineno <= 180]
    assert len(stmts) == 9 and isinstance(stmts[0], ast.Assign) and stmts[0].targets[0].id == "adj", \
        "tensormain.py:170-180 is not the A_hat block any more"
    code = compile(ast.Module(body=stmts, type_ignores=[]), path, "exec")
    sp_shim = types.SimpleNamespace(errstate=np.errstate, sqrt=np.sqrt, isinf=np.isinf, sparse=scipy.sparse)
    nx_shim = types.SimpleNamespace(
        adjacency_matrix=lambda g, nodelist=None, weight="weight":
            scipy.sparse.csr_matrix(nx.adjacency_matrix(g, nodelist=list(nodelist), weight=weight)))
    users = list(range(n))
    ns = {"nx": nx_shim, "sp": sp_shim, "xrange": range, "dl": types.SimpleNamespace(graph=graph),
          "U_train": users, "U_dev": [], "U_test": [], "dtype": "float64", "logging": logging}
    exec(code, ns)
    H64 = scipy.sparse.csr_matrix(ns["H"])
    H32 = scipy.sparse.csr_matrix(H64.astype("float32"))          # tensormain.py:221
    H32.sort_indices()
    return H64, H32


def reference_geo_eval():
    path = os.path.join(REF_DIR, "tensormain.py")
    node = _function_node(path, "geo_eval")
    assert 38 <= node.lineno <= 40
    code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")

    def haversine(p1, p2):
        lat1, lon1, lat2, lon2 = (np.radians(v) for v in (p1[0], p1[1], p2[0], p2[1]))
        d = np.sin((lat2 - lat1) * 0.5) ** 2 + np.cos(lat1) * np.cos(lat2) * np.sin((lon2 - lon1) * 0.5) ** 2
        return 2 * 6371.0088 * np.arcsin(np.sqrt(d))
    ns = {"haversine": haversine, "np": np, "logging": logging}
    exec(code, ns)
    return ns["geo_eval"]


# ----------------------------------------------------------------------------- seeded inputs
def mention_graph(rng, n, avg_deg, isolated=()):
    g = nx.Graph()
    g.add_nodes_from(range(n))
    w = rng.pareto(1.5, size=n) + 1.0
    w /= w.sum()
    m = int(n * avg_deg / 2)
    src = rng.choice(n, size=m, p=w)
    dst = rng.randint(0, n, size=m)
    iso = set(isolated)
    for a, b in zip(src, dst):
        if a != b and a not in iso and b not in iso:
            g.add_edge(int(a), int(b))                 # no 'w' attribute -> binary adjacency
    return g


def random_csr(rng, n, v, mean_nnz, empty_rows=()):
    rows, cols, vals = [], [], []
    for r in range(n):
        if r in empty_rows:
            continue
        k = max(1, rng.poisson(mean_nnz))
        c = np.unique(rng.randint(0, v, size=k))
        rows += [r] * len(c)
        cols += list(c)
        vals += list(rng.rand(len(c)) + 0.05)
    x = scipy.sparse.csr_matrix((np.array(vals, F32), (rows, cols)), shape=(n, v))
    x.sort_indices()
    return x


def main():
    ref = import_reference_layers()
    rng = np.random.RandomState(77)
    out = {}

    # ---- A_hat: tensormain.py:170-180,221 on three graphs (isolated nodes, a clique, power law)
    graphs = {"powerlaw": mention_graph(rng, 300, 8, isolated=(5, 17)),
              "sparse": mention_graph(rng, 120, 2, isolated=(0,))}
    clique = nx.Graph()
    clique.add_nodes_from(range(6))
    clique.add_edges_from([(0, 1), (0, 2), (1, 2), (3, 4)])
    graphs["clique"] = clique
    ahat = {}
    for name, g in graphs.items():
        n = g.number_of_nodes()
        H64, H32 = reference_ahat(g, n)
        e = np.array(sorted(g.edges()), dtype=np.int64).reshape(-1, 2)
        out["ahat_%s__edges" % name] = e
        out["ahat_%s__n" % name] = np.array([n], np.int64)
        out["ahat_%s__indptr" % name] = H32.indptr.astype(np.int32)
        out["ahat_%s__indices" % name] = H32.indices.astype(np.int32)
        out["ahat_%s__data32" % name] = H32.data.astype(F32)
        H64.sort_indices()
        out["ahat_%s__data64" % name] = H64.data.astype(np.float64)
        ahat[name] = H32
        print("A_hat", name, "n", n, "nnz", H32.nnz)

    # ---- layer forwards on the reference classes
    n, V, h, C = 300, 220, 48, 17
    H = ahat["powerlaw"]
    X = random_csr(rng, n, V, 12, empty_rows=(3, 150))
    glorot = lambda a, b: rng.uniform(-np.sqrt(6.0 / (a + b)), np.sqrt(6.0 / (a + b)), size=(a, b)).astype(F32)
    W1, b1 = glorot(V, h), (rng.randn(h) * 0.1).astype(F32)
    W2, b2 = glorot(h, C), (rng.randn(C) * 0.1).astype(F32)
    Wh, bh = glorot(h, h), (rng.randn(h) * 0.1).astype(F32)
    Hin = (rng.randn(n, h) * 0.5).astype(F32)
    idx_dup = rng.choice(n, size=90, replace=True).astype(np.int32)        # tensormain.py:226 samples with replacement
    idx_all = np.arange(n, dtype=np.int32)
    out.update({"X__indptr": X.indptr.astype(np.int32), "X__indices": X.indices.astype(np.int32), "X__data": X.data,
                "X__shape": np.array(X.shape, np.int64), "W1": W1, "b1": b1, "W2": W2, "b2": b2, "Wh": Wh, "bh": bh,
                "Hin": Hin, "idx_dup": idx_dup, "idx_all": idx_all})
    shape_in = (None, V)

    for nl in ("rectify", "tanh", "sigmoid", "identity"):
        ly = ref.SparseInputDenseLayer(shape_in, num_units=h, W=W1, b=b1, nonlinearity=NONLIN[nl])
        out["sid_%s" % nl] = np.asarray(ly.get_output_for(X), dtype=F32)
        ly = ref.SparseConvolutionDenseLayer(shape_in, H=H, num_units=h, W=W1, b=b1, nonlinearity=NONLIN[nl])
        out["scd_%s" % nl] = np.asarray(ly.get_output_for(X), dtype=F32)
    ly = ref.SparseInputDenseLayer(shape_in, num_units=h, W=W1, b=None, nonlinearity=NONLIN["rectify"])
    out["sid_nobias"] = np.asarray(ly.get_output_for(X), dtype=F32)
    for cls in (ref.SparseInputDenseLayer, ref.SparseConvolutionDenseLayer, ref.SparseInputDropoutLayer):
        ly = cls(shape_in, p=0.5) if cls is ref.SparseInputDropoutLayer else \
            cls(shape_in, num_units=h, W=W1, b=b1, nonlinearity=NONLIN["rectify"])
        try:
            ly.get_output_for(np.zeros((n, V), F32))
            raise SystemExit("dense input did not raise")
        except ValueError as e:
            out["error_text"] = np.array(str(e))
    ly = ref.SparseInputDropoutLayer(shape_in, p=0.5)
    assert ly.get_output_for(X, deterministic=True) is X                   # lasagne_layers.py:37-38
    ly0 = ref.SparseInputDropoutLayer(shape_in, p=0)
    assert ly0.get_output_for(X) is X

    for nl, W, b, tag in (("softmax", W2, b2, "out"), ("rectify", Wh, bh, "hid"), ("tanh", Wh, bh, "hidtanh")):
        for iname, idx in (("dup", idx_dup), ("all", idx_all)):
            ly = ref.ConvolutionDenseLayer((None, h), H=H, num_units=W.shape[1], W=W, b=b, nonlinearity=NONLIN[nl])
            out["cd_%s_%s" % (tag, iname)] = np.asarray(ly.get_output_for(Hin, target_indices=idx), dtype=F32)
    ly = ref.ConvolutionDenseLayer((None, h), H=H, num_units=C, W=W2, b=None, nonlinearity=NONLIN["identity"])
    out["cd_logits_nobias_dup"] = np.asarray(ly.get_output_for(Hin, target_indices=idx_dup), dtype=F32)

    # the reference 2-layer GCN forward, composed as mlpconv.py:205-216 composes it (dropout off)
    l1 = ref.SparseConvolutionDenseLayer(shape_in, H=H, num_units=h, W=W1, b=b1, nonlinearity=NONLIN["rectify"])
    a1 = l1.get_output_for(X)
    l2 = ref.ConvolutionDenseLayer((None, h), H=H, num_units=C, W=W2, b=b2, nonlinearity=NONLIN["softmax"])
    out["gcn2_probs_dup"] = np.asarray(l2.get_output_for(a1, target_indices=idx_dup), dtype=F32)

    # ---- geo_eval: tensormain.py:38-54
    geo_eval = reference_geo_eval()
    k, m = 40, 500
    med = np.stack([rng.uniform(25, 49, k), rng.uniform(-124, -67, k)], axis=1)
    true = med[rng.randint(0, k, m)] + rng.randn(m, 2) * 1.2
    pred = rng.randint(0, k, m)
    pred[: m // 2] = np.argmin(((true[: m // 2, None, :] - med[None]) ** 2).sum(-1), axis=1)
    users = ["u%d" % i for i in range(m)]
    user_loc = {u: "%r,%r" % (float(a), float(b)) for u, (a, b) in zip(users, true)}
    lat_med = {str(c): float(med[c, 0]) for c in range(k)}
    lon_med = {str(c): float(med[c, 1]) for c in range(k)}
    mean, median, acc = geo_eval(pred, pred, users, lat_med, lon_med, user_loc)
    out.update({"geo__medians": med, "geo__true": true, "geo__pred": pred.astype(np.int64),
                "geo__result": np.array([mean, median, acc], np.float64)})
    try:
        geo_eval(pred, pred[:-1], users, lat_med, lon_med, user_loc)
        raise SystemExit("length mismatch did not assert")
    except AssertionError as e:
        out["geo__assert_text"] = np.array(str(e))

    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
