"""The oracle itself: known-answer identities stated by the reference code, and a cross-check
of the hand-written backward against torch-CPU autograd (the reference has no tests, so the
GCN oracle is 'parity unpinned'; these are the strongest anchors available -- SURVEY 8c)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import gcn_oracle as go

F32 = np.float32


def small_problem(seed=0, n=120, v=60, h=16, c=7, n_idx=50, dup=False):
    rng = np.random.RandomState(seed)
    adj = sp.random(n, n, density=0.06, random_state=rng, format="csr")
    adj = ((adj + adj.T) > 0).astype(np.float64)
    adj.setdiag(0)
    adj.eliminate_zeros()
    A = go.build_ahat(adj)
    X = sp.random(n, v, density=0.15, random_state=rng, format="csr", dtype=np.float32)
    idx = rng.choice(n, size=n_idx, replace=dup).astype(np.int32)
    y = rng.randint(0, c, size=n_idx).astype(np.int32)
    return rng, X, A, idx, y


# ------------------------------------------------------------------ A_hat
def test_ahat_identities():
    # isolated node -> A_ii = 1 (setdiag then d = 1, tensormain.py:172-177)
    adj = sp.csr_matrix((3, 3))
    A = go.build_ahat(adj).toarray()
    assert np.array_equal(A, np.eye(3, dtype=F32))
    # 3-clique incl. self loops -> every entry 1/3
    adj = sp.csr_matrix(np.ones((3, 3)) - np.eye(3))
    A = go.build_ahat(adj).toarray()
    assert np.allclose(A, 1.0 / 3.0, rtol=0, atol=1e-7)
    # symmetric bit for bit (float64 normalise then cast), so A^T = A serves the backward
    rng = np.random.RandomState(1)
    adj = sp.random(200, 200, density=0.05, random_state=rng, format="csr")
    adj = ((adj + adj.T) > 0).astype(np.float64)
    A = go.build_ahat(adj)
    assert A.dtype == F32 and (A != A.T).nnz == 0
    d = np.asarray(((adj + sp.eye(200)) > 0).sum(axis=1)).ravel()
    i, j = 5, A[5].indices[0]
    assert A[i, j] == F32(1.0 / np.sqrt(d[i]) * 1.0 / np.sqrt(d[j]))


def test_ahat_product_builder_is_bit_identical(built_lib):
    from graphconvgeo_b200.sparse import build_ahat_host
    from graphconvgeo_b200.synth import powerlaw_graph
    for seed in range(3):
        adj = powerlaw_graph(2000 + 100 * seed, 6 + seed, seed)
        a, b = go.build_ahat(adj), build_ahat_host(adj)
        assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
        assert np.array_equal(a.data, b.data)
    # an adjacency that already carries (non-unit) diagonal entries and weights
    rng = np.random.RandomState(3)
    w = sp.random(50, 50, density=0.2, random_state=rng, format="csr")
    w = w + w.T + sp.diags(rng.rand(50))
    a, b = go.build_ahat(w), build_ahat_host(w)
    assert np.array_equal(a.indices, b.indices) and np.array_equal(a.data, b.data)


# ------------------------------------------------------------------ layers
def test_layer_forwards_follow_the_reference_order():
    rng, X, A, idx, y = small_problem()
    W = rng.randn(X.shape[1], 8).astype(F32)
    b = rng.randn(8).astype(F32)
    Xd, Ad = X.toarray(), A.toarray()
    out = go.sparse_input_dense(X, W, b, "tanh")
    assert np.allclose(out, np.tanh(Xd @ W + b), atol=1e-5)
    out = go.sparse_convolution_dense(X, W, b, A, "rectify")
    assert np.allclose(out, np.maximum(Ad @ (Xd @ W) + b, 0), atol=1e-5)     # bias AFTER propagation
    Hin = rng.randn(X.shape[0], 5).astype(F32)
    W2 = rng.randn(5, 8).astype(F32)
    out = go.convolution_dense(Hin, W2, b, A, idx, "softmax")
    ref = (Ad @ (Hin @ W2) + b)[idx]
    ref = np.exp(ref - ref.max(1, keepdims=True))
    ref /= ref.sum(1, keepdims=True)
    assert np.allclose(out, ref, atol=1e-6) and out.shape == (len(idx), 8)
    with pytest.raises(ValueError, match="must be sparse"):
        go.sparse_input_dense(Xd, W, b)
    with pytest.raises(ValueError, match="must be sparse"):
        go.sparse_convolution_dense(Xd, W, b, A)


# ----------------------------------------------------- backward vs autograd
def torch_loss(params, X, A, idx, y, n_layers, highway, c_out, c_hid, act):
    Xt = torch.tensor(X.toarray(), dtype=torch.float64)
    At = torch.tensor(A.toarray(), dtype=torch.float64)
    f = {"rectify": torch.relu, "tanh": torch.tanh}[act]
    it = iter(params)
    W, b = next(it), next(it)
    h = f(At @ (Xt @ W) + b)
    reg = 0.5 * c_hid * (W.abs().sum() + (W * W).sum())
    for _ in range(n_layers - 2):
        W, b = next(it), next(it)
        hc = f(At @ (h @ W) + b)
        reg = reg + 0.5 * c_hid * (W.abs().sum() + (W * W).sum())
        if highway:
            Wg, bg = next(it), next(it)
            g = torch.sigmoid(h @ Wg + bg)
            h = g * hc + (1 - g) * h
            reg = reg + 0.5 * c_hid * (Wg.abs().sum() + (Wg * Wg).sum())
        else:
            h = hc
    W, b = next(it), next(it)
    logits = (At @ (h @ W) + b)[torch.tensor(idx, dtype=torch.long)]
    reg = reg + 0.5 * c_out * (W.abs().sum() + (W * W).sum())
    ce = torch.nn.functional.cross_entropy(logits, torch.tensor(y, dtype=torch.long), reduction="mean")
    return ce + reg


@pytest.mark.parametrize("n_layers,highway,act,dup", [(2, False, "rectify", False), (2, False, "rectify", True),
                                                       (3, True, "rectify", False), (4, True, "tanh", True),
                                                       (3, False, "tanh", False)])
def test_backward_matches_autograd(n_layers, highway, act, dup):
    rng, X, A, idx, y = small_problem(seed=n_layers, dup=dup)
    c_out, c_hid = 1e-3, 2e-3
    params = go.init_params(rng, X.shape[1], 16, 7, n_layers, highway)
    for p in params:                      # non-zero biases so their grads are exercised
        if p.ndim == 1:
            p[...] = rng.randn(*p.shape).astype(F32) * 0.1
    net = go.GCNOracle(X, A, n_layers, highway, (c_out, c_hid), act)
    loss, acc, grads, _ = net.loss_and_grads(params, idx, y)
    tp = [torch.tensor(p, dtype=torch.float64, requires_grad=True) for p in params]
    tl = torch_loss(tp, X, A, idx, y, n_layers, highway, c_out, c_hid, act)
    tl.backward()
    assert abs(float(loss) - float(tl.detach())) < 1e-5
    for g, t in zip(grads, tp):
        ref = t.grad.numpy()
        assert g.dtype == F32 and g.shape == ref.shape
        assert np.all(np.abs(g - ref) <= 2e-6 + 1e-4 * np.abs(ref)), float(np.abs(g - ref).max())


def test_duplicate_indices_accumulate():
    rng, X, A, idx, y = small_problem(seed=9, n_idx=30)
    idx2 = np.concatenate([idx, idx[:10]])
    y2 = np.concatenate([y, y[:10]])
    params = go.init_params(rng, X.shape[1], 16, 7)
    net = go.GCNOracle(X, A)
    _, _, _, c = net.loss_and_grads(params, idx2, y2)
    dP = c["dP_out"]
    # a duplicated node receives two gradient rows, each scaled by 1/len(idx) counting duplicates
    rows = np.flatnonzero(np.abs(dP).sum(1) > 0)
    assert set(rows) == set(idx2.tolist())
    assert np.isclose(np.abs(dP).sum(), np.abs(c["probs"] - np.eye(7, dtype=F32)[y2]).sum() / len(idx2), rtol=1e-4)


# -------------------------------------------------------------------- Adam
def test_adam_is_lasagne_adam():
    rng = np.random.RandomState(0)
    p = [rng.randn(5, 3).astype(F32)]
    p0 = p[0].copy()
    st = go.AdamState(p)
    g1, g2 = rng.randn(5, 3).astype(F32), rng.randn(5, 3).astype(F32)
    go.adam_step(p, [g1], st)
    # first step: m = .1 g, v = .001 g^2, a_1 = lr*sqrt(.001)/.1 -> step = lr*g/(|g| + eps*sqrt(1000)...) ~ lr*sign(g)
    a1 = 4e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    exp = p0 - a1 * (0.1 * g1) / (np.sqrt(0.001 * g1 * g1) + 1e-8)
    assert np.allclose(p[0], exp, rtol=1e-5, atol=1e-7)
    go.adam_step(p, [g2], st)
    assert st.t == 2
    m = 0.9 * 0.1 * g1 + 0.1 * g2
    v = 0.999 * 0.001 * g1 * g1 + 0.001 * g2 * g2
    a2 = 4e-3 * np.sqrt(1 - 0.999 ** 2) / (1 - 0.9 ** 2)
    assert np.allclose(p[0], exp - a2 * m / (np.sqrt(v) + 1e-8), rtol=1e-5, atol=1e-7)


def test_training_reduces_loss():
    rng, X, A, idx, y = small_problem(seed=4, n_idx=100, dup=True)
    params = go.init_params(rng, X.shape[1], 16, 7, 3, True)
    net = go.GCNOracle(X, A, 3, True, (1e-6, 1e-6))
    hist = go.train_epochs(net, params, idx, y, 40)
    assert hist[-1][0] < hist[0][0] - 0.02


def test_sampled_parity_report_bound_and_reference_noise_floor():
    """oracle/sampled_parity._Report: the scaled error is |got - ref| / (1e-6 + 1e-4|ref|); a check that carries the
    reference's own float32 result is also measured against that result's distance from the float64 values."""
    import numpy as np
    from oracle.sampled_parity import _Report
    rep = _Report()
    ref = np.array([0.0, 1.0, -100.0])
    assert rep.add("exact", ref.copy(), ref) == 0.0
    got = ref + np.array([5e-7, 5e-5, 0.0])             # half the absolute term, half the relative term
    assert abs(rep.add("half", got, ref) - 0.5) < 0.01
    r = rep.add("over with floor", ref + np.array([2e-6, 0, 0]), ref, reference_f32=ref + np.array([3e-6, 0, 0]))
    assert abs(r - 2.0) < 1e-9
    s = rep.summary()
    assert s["worst_check"] == "over with floor" and abs(s["max_scaled_err"] - 2.0) < 1e-9
    fl = s["reference_f32_noise"]["over with floor"]
    assert abs(fl["max_scaled_err"] - 3.0) < 1e-9 and abs(fl["gpu_vs_reference_f32"] - 1.0) < 1e-3
    # 2.0 against a floor of 3.0: inside the reference's own noise -> excess < 1; "half" has no floor -> 0.5
    assert abs(s["max_scaled_err_over_reference_noise"] - 2.0 / 3.0) < 1e-9
    rep.add("over without floor", ref + np.array([4e-6, 0, 0]), ref)
    s = rep.summary()
    assert s["worst_check_over_reference_noise"] == "over without floor" and abs(s["max_scaled_err_over_reference_noise"] - 4.0) < 1e-9
    assert s["checks"]["over without floor"]["frac_over"] == 1.0 / 3.0
