"""SURVEY.md section 8(f) row 2 on the GPU: the smoothing SpGEMM X_conv = H * X (main.py:528-530) bit for bit
against scipy's csr_matmat (values, and the canonical entry order astype() leaves), the device-side minibatch slicing / transpose, and the
minibatch MLP (mlp.py:121-314) against the oracle within the north_star tolerance."""
import numpy as np
import pytest
import scipy.sparse as sp

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import gcn_oracle as go  # noqa: E402
from oracle import mlp_oracle as mo  # noqa: E402
from util import assert_close, random_csr  # noqa: E402


def _tfidf_like(rng, n, v, terms):
    x = random_csr(rng, n, v, terms, empty_frac=0.05)
    x.data = np.abs(x.data).astype(np.float32) + np.float32(0.01)
    return x


def _graph(rng, n, deg, hubs=()):
    a = random_csr(rng, n, n, deg, hub_rows=hubs, hub_deg=min(n, 400), empty_frac=0.1)
    a = ((a + a.T) > 0).astype(np.float64)
    return sp.csr_matrix(a)


def _raw(m):
    return m.indptr.astype(np.int64), m.indices.astype(np.int64), m.data


def _dev_raw(c):
    return (c.indptr.cpu().numpy().astype(np.int64), c.indices.cpu().numpy().astype(np.int64), c.data.cpu().numpy())


@pytest.mark.parametrize("n,v,deg,terms,hubs", [(300, 500, 6, 20, ()), (1500, 3000, 8, 40, (3, 700)),
                                                 (64, 70000, 3, 90, ()), (1, 10, 1, 3, ())])
def test_smoothing_spgemm_bit_exact_float64_ahat(n, v, deg, terms, hubs):
    from graphconvgeo_b200.sparse import smooth_features
    rng = np.random.RandomState(n + v)
    H = go.build_ahat(_graph(rng, n, deg, hubs), dtype="float64")        # main.py:513-522 keeps float64
    X = _tfidf_like(rng, n, v, terms)
    ref = mo.smooth_features(H, X)                                       # scipy csr_matmat + astype('float32')
    assert ref.has_sorted_indices                                        # astype() canonicalises
    got = smooth_features(H, X)
    torch.cuda.synchronize()
    assert got.shape == ref.shape
    for g, r, name in zip(_dev_raw(got), _raw(ref), ("indptr", "indices", "data")):
        assert g.shape == r.shape and np.array_equal(g, r), "SpGEMM %s differs from scipy" % name


def test_spgemm_float32_operands_and_unsorted_rows():
    from graphconvgeo_b200.sparse import CSRMatrix, spgemm
    rng = np.random.RandomState(3)
    A = random_csr(rng, 400, 350, 7, hub_rows=(5,), hub_deg=300)
    B = random_csr(rng, 350, 900, 25)
    # shuffle the entries inside A's rows: the kernel must follow the stored order like scipy does
    ip, ix, d = A.indptr, A.indices.copy(), A.data.copy()
    for r in range(A.shape[0]):
        p = rng.permutation(ip[r + 1] - ip[r])
        ix[ip[r]:ip[r + 1]] = ix[ip[r]:ip[r + 1]][p]
        d[ip[r]:ip[r + 1]] = d[ip[r]:ip[r + 1]][p]
    A = sp.csr_matrix((d, ix, ip), shape=A.shape)
    ref = A * B                                                          # float32 * float32: float32 sums
    ref.sort_indices()                                                   # the kernel emits ascending columns
    got = spgemm(CSRMatrix.from_scipy(A, sort_indices=False), CSRMatrix.from_scipy(B))
    torch.cuda.synchronize()
    gi, gx, gd = _dev_raw(got)
    assert np.array_equal(gi, ref.indptr) and np.array_equal(gx, ref.indices)
    # scipy drops entries whose sum is exactly 0.0; random normals never cancel exactly here
    assert np.array_equal(gd, ref.data)


def test_spgemm_edge_cases():
    from graphconvgeo_b200.sparse import CSRMatrix, spgemm
    A = sp.csr_matrix((np.array([1.0, 2.0], np.float32), np.array([0, 2]), np.array([0, 0, 2, 2])), shape=(3, 3))
    B = sp.csr_matrix((np.array([3.0], np.float32), np.array([4]), np.array([0, 0, 0, 1])), shape=(3, 6))
    got = spgemm(CSRMatrix.from_scipy(A), CSRMatrix.from_scipy(B))
    torch.cuda.synchronize()
    ref = A * B
    ref.sort_indices()
    gi, gx, gd = _dev_raw(got)
    assert np.array_equal(gi, ref.indptr) and np.array_equal(gx, ref.indices) and np.array_equal(gd, ref.data)
    with pytest.raises(ValueError):
        spgemm(CSRMatrix.from_scipy(B), CSRMatrix.from_scipy(B))


def test_minibatch_slice_and_transpose_on_device():
    from graphconvgeo_b200.sparse import CSRMatrix
    rng = np.random.RandomState(8)
    X = random_csr(rng, 700, 1200, 15, hub_rows=(9,), hub_deg=600)
    Xd = CSRMatrix.from_scipy(X)
    rows = rng.permutation(700)[:256]
    rows[5] = rows[6]                                                    # duplicates are legal
    b = Xd.gather_rows_device(rows)
    torch.cuda.synchronize()
    ref = X[rows]
    for g, r in zip(_dev_raw(b), _raw(ref)):
        assert np.array_equal(g, r)
    t = b.T                                                              # device transpose (stable)
    torch.cuda.synchronize()
    rt = sp.csr_matrix(ref.T)
    rt.sort_indices()
    for g, r in zip(_dev_raw(t), _raw(rt)):
        assert np.array_equal(g, r)
    assert t.shape == (1200, 256)
    # empty matrix
    e = CSRMatrix.from_scipy(sp.csr_matrix((5, 9), dtype=np.float32)).gather_rows_device(np.array([0, 4]))
    et = e.T
    torch.cuda.synchronize()
    assert et.indptr.cpu().numpy().tolist() == [0] * 10 and et.nnz == 0


def _problem(seed, n_train=600, n_dev=150, v=400, classes=7, terms=12):
    rng = np.random.RandomState(seed)
    n = n_train + n_dev
    X = _tfidf_like(rng, n, v, terms)
    proj = rng.standard_normal((v, classes)).astype(np.float32)
    Y = np.asarray(X @ proj).argmax(-1).astype(np.int32)
    assert len(set(Y[:n_train].tolist())) == classes
    return X[:n_train], Y[:n_train], X[n_train:], Y[n_train:], rng


@pytest.mark.parametrize("add_hidden,act", [(True, "rectify"), (True, "tanh"), (False, "rectify")])
def test_mlp_minibatch_trajectory_matches_oracle(add_hidden, act):
    from graphconvgeo_b200.mlp import MLP
    Xtr, Ytr, Xdv, Ydv, rng = _problem(21)
    hidden, classes, reg = 48, 7, (1e-4, 2e-4)
    params = mo.init_params(rng, Xtr.shape[1], hidden, classes, add_hidden)
    net = mo.MLPOracle(reg, act, add_hidden)
    ref_params = [p.copy() for p in params]
    steps, epochs, best = mo.fit(net, ref_params, Xtr, Ytr, Xdv, Ydv, n_epochs=2, batch_size=128, seed=4)
    clf = MLP(n_epochs=2, batch_size=128, init_parameters=[p.copy() for p in params], add_hidden=add_hidden,
              regul_coefs=list(reg), hidden_layer_size=hidden, drop_out=False, nonlinearity=act, seed=4)
    clf.prepare(Xtr, Ytr)
    got_steps = []
    orig = clf.f_train

    def recording(xb, yb):
        hb = orig(xb, yb)
        got_steps.append(clf.train_results())
        return hb

    clf.f_train = recording
    for _ in range(2):
        assert clf.train_epoch() == 600 // 128                          # mlp.py:86: full batches only
    assert len(got_steps) == len(steps) == 8
    for (lg, ag), (lr_, ar) in zip(got_steps, steps):
        assert abs(lg - lr_) <= 1e-5 + 2e-4 * abs(lr_), (lg, lr_)
        assert abs(ag - ar) <= 1.0 / 128 + 1e-6
    for p_gpu, p in zip(clf.get_param_values(), ref_params):
        assert_close(p_gpu, p, atol=2e-4, rtol=2e-3, what="parameters after 8 Adam steps")
    l_ref, a_ref = net.loss_acc(ref_params, Xdv, Ydv)
    l_gpu, a_gpu = clf.f_val(clf._to_device(Xdv), clf._labels(Ydv))
    assert abs(l_gpu - float(l_ref)) <= 1e-5 + 2e-4 * abs(float(l_ref))
    assert abs(a_gpu - a_ref) <= 2.0 / len(Ydv)


def test_mlp_single_step_gradients_match_oracle():
    from graphconvgeo_b200.mlp import MLP
    Xtr, Ytr, _, _, rng = _problem(33)
    hidden, classes, reg = 32, 7, (1e-4, 2e-4)
    params = mo.init_params(rng, Xtr.shape[1], hidden, classes)
    for p in params:
        if p.ndim == 1:
            p[...] = (rng.standard_normal(p.shape) * 0.05).astype(np.float32)
    net = mo.MLPOracle(reg)
    idx = rng.permutation(Xtr.shape[0])[:200]
    loss, acc, grads = net.loss_and_grads(params, Xtr[idx], Ytr[idx])
    clf = MLP(batch_size=200, init_parameters=[p.copy() for p in params], regul_coefs=list(reg),
              hidden_layer_size=hidden)
    clf.prepare(Xtr, Ytr)
    clf.f_train(clf._batch(clf.Xd_train, idx), clf._labels(Ytr[idx]))
    torch.cuda.synchronize()
    l_gpu, a_gpu = clf.train_results()
    assert abs(l_gpu - float(loss)) <= 1e-6 + 1e-5 * abs(float(loss)) and abs(a_gpu - acc) < 1e-6
    coefs = [reg[1], 0.0, reg[0], 0.0]
    for g, p, c, r in zip(clf.get_grad_values(), params, coefs, grads):
        g = g + np.float32(0.5) * np.float32(c) * (np.sign(p) + np.float32(2) * p)     # reg grad lives in the Adam kernel
        assert_close(g, r, what="MLP gradient")


def test_mlp_fit_api_dense_input_and_dropout():
    from graphconvgeo_b200.mlp import MLP
    Xtr, Ytr, Xdv, Ydv, _ = _problem(5, n_train=800, n_dev=200, classes=5)
    clf = MLP(n_epochs=15, batch_size=100, regul_coefs=[1e-6, 1e-6], hidden_layer_size=64, drop_out=True,
              drop_out_coefs=[0.5, 0.2], early_stopping_max_down=10, seed=1, learning_rate=1e-2)
    clf.fit(Xtr, Ytr, Xdv, Ydv)                                          # main.py:552-555
    acc = clf.accuracy(Xdv, Ydv)
    assert acc == pytest.approx(clf.best_val[1]) and acc > 0.4        # chance is 0.2
    pred = clf.predict(Xdv)
    proba = clf.predict_proba(Xdv)
    assert pred.dtype == np.int64 and pred.shape == (200,) and proba.shape == (200, 5)
    assert np.array_equal(pred, proba.argmax(-1)) and np.allclose(proba.sum(-1), 1.0, atol=1e-5)
    assert abs((pred == Ydv).mean() - acc) < 1e-6 and clf.score(Xdv, Ydv) == acc
    assert clf.get_embedding(Xdv).shape == (200, 64)
    # dense input takes the DenseLayer branch (mlp.py:172-175) and must agree with the sparse one
    dense = MLP(n_epochs=1, batch_size=100, init_parameters=clf.get_param_values(), hidden_layer_size=64)
    dense.prepare(Xtr.toarray(), Ytr)
    assert_close(dense.predict_proba(Xdv.toarray()), proba, atol=1e-6, rtol=1e-4, what="dense vs sparse input")
    with pytest.raises(ValueError):
        MLP(n_epochs=1, batch_size=5000, hidden_layer_size=8).fit(Xtr, Ytr, Xdv, Ydv)


def test_smoothing_then_mlp_pipeline():
    """main.py:509-556 end to end on a planted-community problem: smoothing must help the classifier."""
    from graphconvgeo_b200.mlp import MLP
    from graphconvgeo_b200.sparse import smooth_features
    rng = np.random.RandomState(2)
    n, v, classes = 1200, 300, 4
    comm = rng.randint(0, classes, size=n)
    rows, cols = [], []
    for i in range(n):
        same = np.flatnonzero(comm == comm[i])
        for j in rng.choice(same, size=6):
            rows.append(i); cols.append(j)
    adj = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, n))
    adj = sp.csr_matrix(((adj + adj.T) > 0).astype(np.float64))
    X = _tfidf_like(rng, n, v, 4)
    sig = sp.csr_matrix((np.full(n, 0.3, np.float32), (np.arange(n), comm * 5 + rng.randint(0, 5, size=n))), shape=(n, v))
    keep = sp.diags((rng.rand(n) < 0.3).astype(np.float32))              # only 30 % of the users carry the signal
    X = sp.csr_matrix(X + keep @ sig, dtype=np.float32)
    X.sum_duplicates()
    H = go.build_ahat(adj, dtype="float64")
    Xc = smooth_features(H, X)
    ref = mo.smooth_features(H, X)
    assert np.array_equal(Xc.data.cpu().numpy(), ref.data) and np.array_equal(Xc.indices.cpu().numpy(), ref.indices)
    tr, dv = np.arange(0, 900), np.arange(900, n)

    def run(feats):
        clf = MLP(n_epochs=30, batch_size=100, regul_coefs=[1e-6, 1e-6], hidden_layer_size=32, seed=0, learning_rate=1e-2)
        clf.fit(feats[tr], comm[tr], feats[dv], comm[dv])
        return clf.best_val[1]

    acc_smooth = run(ref)            # scipy copy of the (bit-identical) smoothed features: row slicing on the host
    acc_raw = run(X)
    assert acc_smooth > acc_raw + 0.1 and acc_smooth > 0.8


def test_smoothing_larger_than_int32_offsets_comes_back_as_row_blocks():
    """Twitter-World smoothing has more than 2^31 non-zeros: sparse.spgemm then returns consecutive row blocks
    (RowBlockedCSR).  Exercised here with a small limit: blocks == scipy's product bit for bit, minibatch row
    gathers (any order, duplicates, rows from several blocks) == scipy slicing, and the minibatch MLP trains to the
    same parameters from the blocked matrix as from the single one."""
    from graphconvgeo_b200.mlp import MLP
    from graphconvgeo_b200.sparse import CSRMatrix, RowBlockedCSR, smooth_features
    rng = np.random.RandomState(31)
    n, v = 900, 400
    X = _tfidf_like(rng, n, v, 10)
    H = go.build_ahat(_graph(rng, n, 6, hubs=(5,)), dtype="float64")
    ref = mo.smooth_features(H, X)
    one = smooth_features(H, X)
    assert isinstance(one, CSRMatrix)
    blk = smooth_features(H, X, max_block_nnz=ref.nnz // 5 + 7)
    assert isinstance(blk, RowBlockedCSR) and len(blk.blocks) >= 5 and blk.shape == ref.shape and blk.nnz == ref.nnz
    assert all(b.nnz <= ref.nnz // 5 + 7 for b in blk.blocks)
    got = blk.to_scipy()
    assert np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices) and np.array_equal(got.data, ref.data)
    rows = rng.choice(n, size=257, replace=True)
    g = blk.gather_rows_device(rows)
    want = ref[rows]
    assert np.array_equal(g.indptr.cpu().numpy(), want.indptr) and np.array_equal(g.indices.cpu().numpy(), want.indices)
    assert np.array_equal(g.data.cpu().numpy(), want.data)
    with pytest.raises(IndexError):
        blk.gather_rows_device([n])
    Y = rng.randint(0, 5, size=n).astype(np.int32)

    def train(feats):
        clf = MLP(n_epochs=2, batch_size=128, regul_coefs=[1e-6, 1e-6], hidden_layer_size=16, seed=3)
        clf.fit(feats, Y, ref[:100], Y[:100])
        return clf.get_param_values() if hasattr(clf, "get_param_values") else [p.cpu().numpy() for p in clf.params]
    pa, pb = train(one), train(blk)
    for a, b in zip(pa, pb):
        assert np.array_equal(a, b)
