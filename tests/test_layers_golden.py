"""The oracle (and the CUDA path) against outputs of the REFERENCE'S OWN SOURCE.

tests/golden/layers_golden.npz was produced by tests/golden/make_layers_golden.py, which imports
/root/reference/lasagne_layers.py under stub theano/lasagne modules and executes the bodies of
lasagne_layers.py:20-29, 53-71, 73-89, the A_hat statements tensormain.py:170-180 (+ cast :221) and
geo_eval (tensormain.py:38-54).  This pins oracle/gcn_oracle.py's forward restatement, its A_hat builder
and oracle/kdtree_oracle.geo_eval to reference code instead of to our reading of it.

CPU tests: oracle == golden (bit-exact: same scipy / BLAS routines in the same order).
GPU tests (-m gpu): the product's layer classes == golden (bit-exact where the layer is sparse products +
bias + relu only; north_star tolerance 1e-6 + 1e-4*|ref| where a dense GEMM or a transcendental is involved).
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import gcn_oracle as go
from oracle import kdtree_oracle as ko
from util import assert_close

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "layers_golden.npz")
NLS = ["rectify", "tanh", "sigmoid", "identity"]


@pytest.fixture(scope="module")
def z():
    return np.load(GOLDEN)


def _ahat(z, name="powerlaw"):
    n = int(z["ahat_%s__n" % name][0])
    return sp.csr_matrix((z["ahat_%s__data32" % name], z["ahat_%s__indices" % name], z["ahat_%s__indptr" % name]),
                         shape=(n, n))


def _adj(z, name):
    n = int(z["ahat_%s__n" % name][0])
    e = z["ahat_%s__edges" % name]
    a = sp.csr_matrix((np.ones(len(e)), (e[:, 0], e[:, 1])), shape=(n, n))
    a = sp.csr_matrix(((a + a.T) > 0).astype(np.float64))
    return a, n


def _X(z):
    return sp.csr_matrix((z["X__data"], z["X__indices"], z["X__indptr"]), shape=tuple(z["X__shape"]))


# ------------------------------------------------------------------------------------------- CPU: oracle
@pytest.mark.parametrize("name", ["powerlaw", "sparse", "clique"])
def test_oracle_ahat_equals_reference_statements(z, name):
    adj, n = _adj(z, name)
    h = go.build_ahat(adj)                               # float32, as after tensormain.py:221
    assert np.array_equal(h.indptr, z["ahat_%s__indptr" % name])
    assert np.array_equal(h.indices, z["ahat_%s__indices" % name])
    assert np.array_equal(h.data, z["ahat_%s__data32" % name])          # bit-exact
    h64 = go.build_ahat(adj, dtype="float64")
    assert np.array_equal(h64.data, z["ahat_%s__data64" % name])


@pytest.mark.parametrize("name", ["powerlaw", "sparse", "clique"])
def test_host_ahat_builder_equals_reference_statements(z, name, built_lib):
    from graphconvgeo_b200.sparse import build_ahat_host
    adj, n = _adj(z, name)
    h = build_ahat_host(adj)
    assert np.array_equal(h.indptr, z["ahat_%s__indptr" % name])
    assert np.array_equal(h.indices, z["ahat_%s__indices" % name])
    assert np.array_equal(h.data, z["ahat_%s__data32" % name])


def test_reference_ahat_identities(z):
    """The identities SURVEY 8(c) lists, read off the reference's own output: symmetric, isolated node -> 1,
    k-clique -> 1/k."""
    h = _ahat(z, "clique").toarray()
    assert np.array_equal(h, h.T)
    assert np.allclose(h[:3, :3], 1.0 / 3) and np.allclose(h[3:5, 3:5], 0.5) and h[5, 5] == 1.0
    hp = _ahat(z, "powerlaw")
    assert (hp != hp.T).nnz == 0
    assert hp[5, 5] == 1.0 and hp[17, 17] == 1.0          # the two isolated nodes


@pytest.mark.parametrize("nl", NLS)
def test_oracle_sparse_layers_equal_reference_bodies(z, nl):
    X, H = _X(z), _ahat(z)
    assert np.array_equal(go.sparse_input_dense(X, z["W1"], z["b1"], nl), z["sid_%s" % nl])            # :20-29
    assert np.array_equal(go.sparse_convolution_dense(X, z["W1"], z["b1"], H, nl), z["scd_%s" % nl])   # :53-71


def test_oracle_sparse_layer_without_bias_and_error_text(z):
    X = _X(z)
    assert np.array_equal(go.sparse_input_dense(X, z["W1"], None, "rectify"), z["sid_nobias"])
    with pytest.raises(ValueError) as e:
        go.sparse_input_dense(np.zeros(X.shape, np.float32), z["W1"], z["b1"])
    assert str(e.value) == str(z["error_text"]) == "Input for this layer must be sparse"


@pytest.mark.parametrize("tag,nl,w,b", [("out", "softmax", "W2", "b2"), ("hid", "rectify", "Wh", "bh"),
                                       ("hidtanh", "tanh", "Wh", "bh")])
@pytest.mark.parametrize("iname", ["dup", "all"])
def test_oracle_convolution_layer_equals_reference_body(z, tag, nl, w, b, iname):
    H = _ahat(z)
    got = go.convolution_dense(z["Hin"], z[w], z[b], H, z["idx_" + iname], nl)                         # :73-89
    assert np.array_equal(got, z["cd_%s_%s" % (tag, iname)])


def test_oracle_network_forward_equals_reference_composition(z):
    """GCNOracle.forward (the 2-layer net of mlpconv.py:205-216) == the reference classes chained."""
    X, H = _X(z), _ahat(z)
    net = go.GCNOracle(X, H, n_layers=2, highway=False)
    params = [z["W1"], z["b1"], z["W2"], z["b2"]]
    c = net.forward(params, z["idx_dup"])
    assert np.array_equal(c["probs"], z["gcn2_probs_dup"])
    ly = go.convolution_dense(z["Hin"], z["W2"], None, H, z["idx_dup"], "identity")
    assert np.array_equal(ly, z["cd_logits_nobias_dup"])


def test_oracle_geo_eval_equals_reference_function(z):
    mean, median, acc = ko.geo_eval(z["geo__true"], z["geo__pred"], z["geo__medians"])
    rmean, rmedian, racc = z["geo__result"]
    assert abs(mean - rmean) < 1e-9 and abs(median - rmedian) < 1e-9 and acc == racc
    assert str(z["geo__assert_text"]).startswith("#preds")


# ------------------------------------------------------------------------------------------- GPU: product
def _dev_layers():
    torch = pytest.importorskip("torch")
    from graphconvgeo_b200 import lasagne_layers as L
    from graphconvgeo_b200.sparse import CSRMatrix
    return torch, L, CSRMatrix


@pytest.mark.gpu
@pytest.mark.parametrize("nl", NLS)
def test_gpu_sparse_layers_equal_reference_bodies(z, nl):
    torch, L, CSRMatrix = _dev_layers()
    X, H = CSRMatrix.from_scipy(_X(z)), CSRMatrix.from_scipy(_ahat(z))
    n, V = X.shape
    l_in = L.InputLayer((None, V), input_var=X)
    sid = L.SparseInputDenseLayer(l_in, num_units=z["W1"].shape[1], W=z["W1"], b=z["b1"], nonlinearity=nl)
    scd = L.SparseConvolutionDenseLayer(l_in, H=H, num_units=z["W1"].shape[1], W=z["W1"], b=z["b1"], nonlinearity=nl)
    a = sid.get_output_for(X).cpu().numpy()
    c = scd.get_output_for(X).cpu().numpy()
    if nl in ("rectify", "identity"):          # sparse products in CSR order + bias + max: bit-exact
        assert np.array_equal(a, z["sid_%s" % nl]) and np.array_equal(c, z["scd_%s" % nl])
    else:                                      # tanhf / expf differ from NumPy's by ulps
        assert_close(a, z["sid_%s" % nl], what="SparseInputDenseLayer " + nl)
        assert_close(c, z["scd_%s" % nl], what="SparseConvolutionDenseLayer " + nl)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,nl,w,b", [("out", "softmax", "W2", "b2"), ("hid", "rectify", "Wh", "bh"),
                                       ("hidtanh", "tanh", "Wh", "bh")])
@pytest.mark.parametrize("iname", ["dup", "all"])
@pytest.mark.parametrize("propagate_first", [False, True])
def test_gpu_convolution_layer_equals_reference_body(z, tag, nl, w, b, iname, propagate_first):
    torch, L, CSRMatrix = _dev_layers()
    from util import to_dev
    H = CSRMatrix.from_scipy(_ahat(z))
    hin = to_dev(z["Hin"])
    l_in = L.InputLayer((None, hin.shape[1]))
    ly = L.ConvolutionDenseLayer(l_in, H=H, num_units=z[w].shape[1], W=z[w], b=z[b], nonlinearity=nl,
                                 propagate_first=propagate_first)
    got = ly.get_output_for(hin, target_indices=z["idx_" + iname]).cpu().numpy()
    assert_close(got, z["cd_%s_%s" % (tag, iname)], what="ConvolutionDenseLayer %s %s" % (tag, iname))
    if nl == "softmax":
        assert np.array_equal(got.argmax(-1), z["cd_%s_%s" % (tag, iname)].argmax(-1))


@pytest.mark.gpu
def test_gpu_ahat_builder_equals_reference_statements(z):
    torch, L, CSRMatrix = _dev_layers()
    from graphconvgeo_b200.sparse import build_ahat_device
    for name in ("powerlaw", "sparse", "clique"):
        adj, n = _adj(z, name)
        adj.sort_indices()
        ip = torch.from_numpy(adj.indptr.astype(np.int32)).cuda()
        ix = torch.from_numpy(adj.indices.astype(np.int32)).cuda()
        h = build_ahat_device(ip, ix, n)
        assert np.array_equal(h.indptr.cpu().numpy(), z["ahat_%s__indptr" % name])
        assert np.array_equal(h.indices.cpu().numpy(), z["ahat_%s__indices" % name])
        assert np.array_equal(h.data.cpu().numpy(), z["ahat_%s__data32" % name])


@pytest.mark.gpu
def test_gpu_network_predictions_equal_reference_composition(z):
    """MLPCONV (2 layers, no gate -- the reference network) with the golden parameters: probabilities within
    tolerance of, and argmax identical to, the reference classes chained (mlpconv.py:205-216, :223)."""
    torch, L, CSRMatrix = _dev_layers()
    from graphconvgeo_b200.mlpconv import MLPCONV
    X, H = _X(z), _ahat(z)
    n = X.shape[0]
    C = z["W2"].shape[1]
    Y = np.arange(n) % C
    m = MLPCONV(hidden_layer_size=z["W1"].shape[1], init_parameters=[z["W1"], z["b1"], z["W2"], z["b2"]],
                regul_coefs=[1e-6, 1e-6], reorder=None)
    idx = z["idx_dup"]
    m.prepare(X, idx, idx, idx, Y, H)
    probs = m.predict_proba("train")
    assert_close(probs, z["gcn2_probs_dup"], what="MLPCONV.predict_proba")
    assert np.array_equal(m.predict("train"), z["gcn2_probs_dup"].argmax(-1))
