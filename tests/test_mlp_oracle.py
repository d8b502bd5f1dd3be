"""CPU checks of the smoothing + minibatch-MLP oracle (oracle/mlp_oracle.py):
its backward against torch autograd, the minibatch rule of mlp.py:81-91, and a plain-Python model of the
row-wise algorithm the CUDA SpGEMM implements against scipy's csr_matmat (values and entry order)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import gcn_oracle as go
from oracle import mlp_oracle as mo
from util import random_csr

F32 = np.float32


def _problem(seed=0, n=90, v=50, c=5):
    rng = np.random.RandomState(seed)
    X = random_csr(rng, n, v, 6, empty_frac=0.05)
    y = rng.randint(0, c, size=n).astype(np.int32)
    return rng, X, y


@pytest.mark.parametrize("add_hidden,act", [(True, "rectify"), (True, "tanh"), (True, "sigmoid"), (False, "rectify")])
def test_mlp_backward_matches_autograd(add_hidden, act):
    rng, X, y = _problem()
    params = mo.init_params(rng, X.shape[1], 12, 5, add_hidden)
    for p in params:
        if p.ndim == 1:
            p[...] = rng.randn(*p.shape).astype(F32) * F32(0.1)
    c_out, c_hid = 3e-3, 7e-3
    net = mo.MLPOracle((c_out, c_hid), act, add_hidden)
    loss, acc, grads = net.loss_and_grads(params, X, y)
    tp = [torch.tensor(p, dtype=torch.float64, requires_grad=True) for p in params]
    Xt = torch.tensor(X.toarray(), dtype=torch.float64)
    pen = lambda W, c: 0.5 * c * (W.abs().sum() + (W * W).sum())
    if add_hidden:
        f = {"rectify": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid}[act]
        logits = f(Xt @ tp[0] + tp[1]) @ tp[2] + tp[3]
        reg = pen(tp[2], c_out) + pen(tp[0], c_hid)                  # mlp.py:222-229
    else:
        logits = Xt @ tp[0] + tp[1]
        reg = pen(tp[0], c_out)
    ref = torch.nn.functional.cross_entropy(logits, torch.tensor(y, dtype=torch.long)) + reg
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-5
    for g, t in zip(grads, tp):
        assert g.dtype == F32 and np.allclose(g, t.grad.numpy(), atol=2e-6, rtol=1e-4)
    l2, a2 = net.loss_acc(params, X, y)
    assert abs(float(l2) - float(loss)) < 1e-6 and a2 == acc
    assert np.allclose(net.predict_proba(params, X).sum(1), 1, atol=1e-5)


def test_minibatches_are_full_batches_of_a_seeded_shuffle():
    rng = np.random.RandomState(3)
    batches = list(mo.iterate_minibatches(23, 5, rng))
    assert len(batches) == 4 and all(len(b) == 5 for b in batches)      # mlp.py:86 drops the ragged tail
    flat = np.concatenate(batches)
    assert len(set(flat.tolist())) == 20
    expect = np.arange(23)
    np.random.RandomState(3).shuffle(expect)
    assert np.array_equal(flat, expect[:20])
    assert list(mo.iterate_minibatches(4, 5, rng)) == []


def test_mlp_fit_learns_and_selects_by_dev_accuracy():
    rng = np.random.RandomState(1)
    X = random_csr(rng, 500, 80, 8, empty_frac=0.0)
    X.data = np.abs(X.data)
    y = np.asarray(X @ rng.randn(80, 4).astype(F32)).argmax(-1).astype(np.int32)
    params = mo.init_params(rng, 80, 24, 4)
    net = mo.MLPOracle((1e-6, 1e-6))
    steps, epochs, best = mo.fit(net, params, X[:400], y[:400], X[400:], y[400:], n_epochs=25, batch_size=50,
                                 seed=0, lr=1e-2)
    assert len(steps) == 25 * 8 and steps[-1][0] < steps[0][0]
    accs = [e[1] for e in epochs]
    assert max(accs) > 0.5
    _, acc_best = net.loss_acc(best, X[400:], y[400:])
    assert acc_best == max(accs)                                         # mlp.py:273-277
    # early stopping (mlp.py:279-284): no improvement allowed -> stops after the first non-improving epoch
    s2, e2, _ = mo.fit(net, mo.init_params(rng, 80, 24, 4), X[:400], y[:400], X[400:], y[400:], 50, 50, 0,
                       lr=0.0, early_stopping_max_down=2)
    assert len(e2) == 4


def spgemm_rowwise_model(A, B, acc_dtype):
    """What gcg_spgemm_{count,fill} compute, in plain Python: per output row, A's entries in stored order,
    B's row entries in stored order, `sums[k] += a*b` from 0 in ``acc_dtype``, columns emitted ascending."""
    indptr, indices, data = [0], [], []
    for i in range(A.shape[0]):
        acc = {}
        for p in range(A.indptr[i], A.indptr[i + 1]):
            j, av = A.indices[p], acc_dtype(A.data[p])
            for q in range(B.indptr[j], B.indptr[j + 1]):
                k = B.indices[q]
                acc[k] = acc_dtype(acc.get(k, acc_dtype(0)) + acc_dtype(av * acc_dtype(B.data[q])))
        for k in sorted(acc):
            indices.append(k)
            data.append(F32(acc[k]))
        indptr.append(len(indices))
    return np.array(indptr), np.array(indices), np.array(data, dtype=F32)


@pytest.mark.parametrize("a_dtype", [np.float64, np.float32])
def test_spgemm_algorithm_model_is_scipy_csr_matmat(a_dtype):
    rng = np.random.RandomState(4)
    A = random_csr(rng, 60, 60, 5, hub_rows=(2,), hub_deg=40, dtype=a_dtype)
    A.data = np.abs(A.data) + a_dtype(0.1)
    ip_, ix_, d_ = A.indptr, A.indices.copy(), A.data.copy()
    for r in range(A.shape[0]):                                          # stored order inside A's rows is what counts
        perm = rng.permutation(ip_[r + 1] - ip_[r])
        ix_[ip_[r]:ip_[r + 1]] = ix_[ip_[r]:ip_[r + 1]][perm]
        d_[ip_[r]:ip_[r + 1]] = d_[ip_[r]:ip_[r + 1]][perm]
    A = sp.csr_matrix((d_, ix_, ip_), shape=A.shape)
    B = random_csr(rng, 60, 45, 7)
    B.data = np.abs(B.data) + F32(0.1)
    raw = (A * B).tocsr()
    assert raw.dtype == a_dtype and not raw.has_sorted_indices           # csr_matmat emits its linked list
    ref = raw.copy()
    ref.sort_indices()
    ip, ix, d = spgemm_rowwise_model(A, B, a_dtype)
    assert np.array_equal(ip, ref.indptr) and np.array_equal(ix, ref.indices)
    assert np.array_equal(d, ref.data.astype(F32))
    if a_dtype == np.float64:
        sm = mo.smooth_features(A, B)                                    # main.py:528-530
        assert sm.dtype == F32 and sm.has_sorted_indices                 # astype() canonicalises the result
        assert np.array_equal(sm.data, d) and np.array_equal(sm.indices, ix)


def test_smoothing_with_reference_normalisation_keeps_row_mass():
    """A_hat rows of a k-clique (self-loops included) are all 1/k (tensormain.py:172-180): smoothing
    with it averages the members' features."""
    adj = sp.csr_matrix(np.ones((3, 3)) - np.eye(3))
    H = go.build_ahat(adj, dtype="float64")
    X = sp.csr_matrix(np.array([[3, 0], [0, 6], [3, 0]], dtype=F32))
    out = mo.smooth_features(H, X).toarray()
    assert out.dtype == F32 and np.allclose(out, [[2, 2]] * 3)
