"""The reference's layer API (lasagne_layers.py:20-89) on the GPU vs the oracle restatement."""
import numpy as np
import pytest
import scipy.sparse as sp

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import gcn_oracle as go  # noqa: E402
from util import assert_close, random_csr, to_dev  # noqa: E402


def problem(seed=0, n=400, v=150, h=48):
    rng = np.random.RandomState(seed)
    adj = sp.random(n, n, density=0.03, random_state=rng, format="csr")
    adj = ((adj + adj.T) > 0).astype(np.float64)
    A = go.build_ahat(adj)
    X = sp.random(n, v, density=0.1, random_state=rng, format="csr", dtype=np.float32)
    return rng, X, A


@pytest.mark.parametrize("act", ["rectify", "tanh", "identity"])
def test_sparse_input_dense_layer(act):
    from graphconvgeo_b200 import lasagne_layers as L
    rng, X, A = problem()
    W = go.glorot_uniform(rng, X.shape[1], 48)
    b = (rng.standard_normal(48) * 0.1).astype(np.float32)
    l_in = L.InputLayer((None, X.shape[1]))
    ly = L.SparseInputDenseLayer(l_in, num_units=48, W=W, b=b, nonlinearity=act)
    assert ly.output_shape == (None, 48)
    out = ly.get_output_for(X).cpu().numpy()
    assert_close(out, go.sparse_input_dense(X, W, b, act))
    with pytest.raises(ValueError, match="Input for this layer must be sparse"):
        ly.get_output_for(torch.zeros(4, X.shape[1], device="cuda"))
    # backward: dW = X^T.(dOut*act'), db = colsum
    dO = rng.standard_normal(out.shape).astype(np.float32)
    ly.backward(to_dev(dO))
    dP = dO * {"rectify": out > 0, "tanh": 1 - out * out, "identity": 1.0}[act]
    assert_close(ly.grads["W"].cpu().numpy(), np.asarray(X.T @ dP.astype(np.float32)), atol=2e-6)
    assert_close(ly.grads["b"].cpu().numpy(), dP.sum(0), atol=2e-5)


def test_sparse_convolution_dense_layer():
    from graphconvgeo_b200 import lasagne_layers as L
    rng, X, A = problem(1)
    W = go.glorot_uniform(rng, X.shape[1], 48)
    b = (rng.standard_normal(48) * 0.1).astype(np.float32)
    ly = L.SparseConvolutionDenseLayer(L.InputLayer((None, X.shape[1])), H=A, num_units=48, W=W, b=b,
                                       nonlinearity=L.nonlinearities.rectify)
    out = ly.get_output_for(X)
    assert_close(out.cpu().numpy(), go.sparse_convolution_dense(X, W, b, A, "rectify"))
    with pytest.raises(ValueError, match="must be sparse"):
        ly.get_output_for(torch.zeros(X.shape[0], X.shape[1], device="cuda"))
    assert [tuple(p.shape) for p in ly.get_params()] == [(X.shape[1], 48), (48,)]
    assert len(ly.get_params(regularizable=True)) == 1          # H is not a parameter (:57,78)


@pytest.mark.parametrize("act", ["softmax", "rectify"])
def test_convolution_dense_layer_with_target_indices(act):
    from graphconvgeo_b200 import lasagne_layers as L
    rng, X, A = problem(2)
    n = X.shape[0]
    Hin = (rng.standard_normal((n, 48)) * 0.3).astype(np.float32)
    W = go.glorot_uniform(rng, 48, 11)
    b = (rng.standard_normal(11) * 0.1).astype(np.float32)
    idx = rng.choice(n, size=123, replace=True).astype(np.int32)
    ly = L.ConvolutionDenseLayer(L.InputLayer((None, 48)), H=A, num_units=11, W=W, b=b, nonlinearity=act)
    out = ly.get_output_for(to_dev(Hin), target_indices=idx)
    ref = go.convolution_dense(Hin, W, b, A, idx, act)
    assert out.shape == (123, 11)
    assert_close(out.cpu().numpy(), ref, atol=2e-6)
    full = ly.get_output_for(to_dev(Hin), target_indices=None) if act != "softmax" else None
    if full is not None:
        assert_close(full.cpu().numpy(), go.convolution_dense(Hin, W, b, A, None, act), atol=2e-6)


def test_highway_layer_forward_backward():
    from graphconvgeo_b200 import lasagne_layers as L
    rng, X, A = problem(3)
    n, h = X.shape[0], 48
    Hin = (rng.standard_normal((n, h)) * 0.3).astype(np.float32)
    W, Wg = go.glorot_uniform(rng, h, h), go.glorot_uniform(rng, h, h)
    b, bg = (rng.standard_normal(h) * 0.1).astype(np.float32), (rng.standard_normal(h) * 0.1).astype(np.float32)
    ly = L.GraphConvLayer(L.InputLayer((None, h)), H=A, highway=True, num_units=h, W=W, b=b, Wg=Wg, bg=bg,
                          nonlinearity="rectify")
    assert isinstance(ly, L.HighwayConvolutionDenseLayer)
    out = ly.get_output_for(to_dev(Hin), train=True)
    hc = go.convolution_dense(Hin, W, b, A, None, "rectify")
    ref, g = go.highway_mix(Hin, hc, Wg, bg)
    assert_close(out.cpu().numpy(), ref, atol=2e-6)
    dO = (rng.standard_normal((n, h)) * 0.1).astype(np.float32)
    dIn = ly.backward(to_dev(dO)).cpu().numpy()
    # oracle backward of the same layer
    dHc = g * dO
    dGp = dO * (hc - Hin) * g * (1 - g)
    dP = dHc * (hc > 0)
    dZ = np.asarray(A @ dP.astype(np.float32))
    ref_dIn = (1 - g) * dO + dZ @ W.T + dGp @ Wg.T
    assert_close(dIn, ref_dIn, atol=3e-6)
    assert_close(ly.grads["W"].cpu().numpy(), Hin.T @ dZ, atol=3e-6)
    assert_close(ly.grads["Wg"].cpu().numpy(), Hin.T @ dGp, atol=3e-6)
    assert_close(ly.grads["b"].cpu().numpy(), dP.sum(0), atol=2e-5)
    assert_close(ly.grads["bg"].cpu().numpy(), dGp.sum(0), atol=2e-5)
    # eval mode does not store H'
    out2 = ly.get_output_for(to_dev(Hin))
    assert ly._Hc is None and torch.equal(out, out2)


def test_sparse_dropout_layer_deterministic_and_errors():
    from graphconvgeo_b200 import lasagne_layers as L
    rng, X, A = problem(4)
    ly = L.SparseInputDropoutLayer(L.InputLayer((None, X.shape[1])), p=0.5)
    assert ly.get_output_for(X, deterministic=True) is X           # lasagne_layers.py:37-38
    with pytest.raises(ValueError, match="must be sparse"):
        ly.get_output_for(np.zeros((3, 3)))
    out = ly.get_output_for(X)
    kept = (out.data != 0).float().mean().item()
    assert 0.35 < kept < 0.65
    nz = out.data[out.data != 0].cpu().numpy()
    src = torch.from_numpy(X.data).cuda()[out.data != 0].cpu().numpy()
    assert_close(nz, src * 2.0)                                     # rescale 1/(1-p)


def test_get_output_and_param_helpers():
    from graphconvgeo_b200 import lasagne_layers as L
    rng, X, A = problem(5)
    l_in = L.InputLayer((None, X.shape[1]), input_var=None)
    l1 = L.SparseConvolutionDenseLayer(l_in, H=A, num_units=32, nonlinearity="rectify", rng=np.random.RandomState(1))
    l2 = L.ConvolutionDenseLayer(l1, H=l1.H, num_units=9, nonlinearity=L.nonlinearities.softmax, rng=np.random.RandomState(2))
    idx = np.arange(50, dtype=np.int32)
    probs = L.get_output(l2, X, target_indices=idx, deterministic=True)
    vals = L.get_all_param_values(l2)
    assert [v.shape for v in vals] == [(X.shape[1], 32), (32,), (32, 9), (9,)]
    ref = go.GCNOracle(X, A).predict_proba(vals, idx)
    assert_close(probs.cpu().numpy(), ref, atol=1e-6)
    a = np.sqrt(6.0 / (X.shape[1] + 32))
    assert np.abs(vals[0]).max() <= a and vals[1].sum() == 0      # GlorotUniform / zero bias
    vals[1][:] = 1.0
    L.set_all_param_values(l2, vals)
    assert torch.all(l1.b == 1.0)
    with pytest.raises(ValueError, match="mismatch"):
        L.set_all_param_values(l2, vals[:-1])


@pytest.mark.parametrize("act", ["softmax", "rectify"])
def test_propagate_first_association_matches_reference_order(act):
    """(H.input).W == H.(input.W) up to float32 rounding, forward and backward (num_units > num_inputs)."""
    from graphconvgeo_b200 import lasagne_layers as L
    rng, X, A = problem(6)
    n, fin, fout = X.shape[0], 24, 40
    Hin = (rng.standard_normal((n, fin)) * 0.3).astype(np.float32)
    W = go.glorot_uniform(rng, fin, fout)
    b = (rng.standard_normal(fout) * 0.1).astype(np.float32)
    idx = rng.choice(n, size=150, replace=True).astype(np.int32)
    outs, grads_w, grads_b, grads_in = [], [], [], []
    dO = (rng.standard_normal((150, fout)) * 0.1).astype(np.float32)
    prev = np.maximum(rng.standard_normal((n, fin)), 0).astype(np.float32)
    for pf in (False, True):
        ly = L.ConvolutionDenseLayer(L.InputLayer((None, fin)), H=A, num_units=fout, W=W, b=b, nonlinearity=act,
                                     propagate_first=pf)
        assert ly.propagate_first == pf
        out = ly.get_output_for(to_dev(Hin), target_indices=idx, logits=True)
        outs.append(out.cpu().numpy().copy())
        dIn = ly.backward(to_dev(dO), input_mask=(to_dev(prev), "rectify"))
        grads_w.append(ly.grads["W"].cpu().numpy().copy())
        grads_b.append(ly.grads["b"].cpu().numpy().copy())
        grads_in.append(dIn.cpu().numpy().copy())
    ref = go.convolution_dense(Hin, W, b, A, idx, "identity" if act == "softmax" else act)
    for o in outs:
        assert_close(o, ref, atol=2e-6)
    assert_close(grads_w[1], grads_w[0], atol=2e-6)
    assert_close(grads_b[1], grads_b[0], atol=2e-6)
    assert_close(grads_in[1], grads_in[0], atol=2e-6)
    auto = L.ConvolutionDenseLayer(L.InputLayer((None, fin)), H=A, num_units=fout)
    assert auto.propagate_first and not L.ConvolutionDenseLayer(L.InputLayer((None, fout)), H=A, num_units=fin).propagate_first


@pytest.mark.parametrize("n,deg,seed", [(1, 0, 0), (50, 3, 1), (3000, 8, 2), (20000, 20, 3)])
def test_device_ahat_builder_is_bit_identical_to_host_and_oracle(n, deg, seed):
    """A_hat = D^-1/2 (A, unit diagonal) D^-1/2 built on the GPU, float64 then cast."""
    from graphconvgeo_b200.sparse import build_ahat_device, build_ahat_host
    from graphconvgeo_b200.synth import powerlaw_graph
    adj = powerlaw_graph(n, deg, seed) if n > 1 else sp.csr_matrix((1, 1))
    if n > 100:                                   # some rows that already carry their diagonal
        adj = sp.csr_matrix(adj + sp.diags((np.arange(n) % 7 == 0).astype(np.float64)))
        adj.data[:] = 1.0
    adj.sort_indices()
    ref = go.build_ahat(adj)
    host = build_ahat_host(adj)
    ip = torch.from_numpy(adj.indptr.astype(np.int32)).cuda()
    ix = torch.from_numpy(adj.indices.astype(np.int32)).cuda()
    got = build_ahat_device(ip, ix, n)
    assert np.array_equal(got.indptr.cpu().numpy(), ref.indptr)
    assert np.array_equal(got.indices.cpu().numpy(), ref.indices)
    assert np.array_equal(got.data.cpu().numpy(), ref.data)
    assert np.array_equal(host.data, ref.data)


@pytest.mark.parametrize("k_head,engine", [(32, "fma"), (64, "tf32x3")])
def test_head_split_products_match_single_pass(k_head, engine):
    """sparse.HeadSplit (dense block of the most frequent terms on the GEMM engine + sparse tail) against the
    single-pass scipy products X.W and X^T.dZ: same values up to float32 re-association, north_star bound."""
    import scipy.sparse as sp
    from graphconvgeo_b200 import ops, synth
    from graphconvgeo_b200.sparse import CSRMatrix, HeadSplit
    rng = np.random.RandomState(12)
    n, V, F = 3000, 2000, 72
    X = synth.tfidf_matrix(n, V, 40, seed=3)
    Xd = CSRMatrix.from_scipy(X)
    hs = HeadSplit(Xd, k_head=k_head)
    assert 0.1 < hs.head_fraction < 0.9 and hs.tail.nnz + hs.head_nnz == X.nnz
    # the split is a partition of X
    rebuilt = hs.tail.to_scipy().toarray()
    rebuilt[:, hs.top.cpu().numpy()] += hs.Xh.cpu().numpy()
    assert np.array_equal(rebuilt, X.toarray())
    W = (rng.standard_normal((V, F)) * 0.1).astype(np.float32)
    b = (rng.standard_normal(F) * 0.1).astype(np.float32)
    dZ = rng.standard_normal((n, F)).astype(np.float32)
    ops.set_gemm_mode(engine)
    try:
        out = ops.alloc_mat(n, F, "cuda")
        hs.product(to_dev(W), out)
        assert_close(out.cpu().numpy(), np.asarray(X @ W, dtype=np.float32), what="X.W")
        hs.product(to_dev(W), out, bias=to_dev(b), act="tanh")
        assert_close(out.cpu().numpy(), np.tanh(np.asarray(X @ W, dtype=np.float32) + b), what="tanh(X.W+b)")
        dW = torch.zeros(V, F, device="cuda")
        ops.spmm(hs.tail.T, to_dev(dZ), out=dW)
        hs.transpose_product_head(to_dev(dZ), dW)
        assert_close(dW.cpu().numpy(), np.asarray(sp.csr_matrix(X.T) @ dZ, dtype=np.float32), atol=2e-5, what="X^T.dZ")
    finally:
        ops.set_gemm_mode("auto")
