"""The oracle's (and the CUDA path's) training loop against a run of the REFERENCE'S OWN ``MLPCONV.fit``.

tests/golden/fit_golden.npz was produced by tests/golden/make_fit_golden.py: /root/reference/mlpconv.py imported
as a module and its ``fit`` (:152-318), ``predict`` and ``predict_proba`` executed on a lazy stub of Theano / Lasagne
(expression nodes evaluated with torch-CPU float32, autograd standing in for theano.grad).  This pins what the
forward-only fixtures cannot: the loss / elastic-net composition with its (out, hid) coefficient order, that f_train
reports the pre-update loss, Adam on [W1, b1, W2, b2], duplicates in train_indices, validation every 10 epochs,
best-snapshot / early-stopping / restore semantics, the pickle name.
Float32 sums are ordered differently by torch, NumPy and the GPU, so values are compared within tolerances."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import gcn_oracle as go
from util import assert_close

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fit_golden.npz")


@pytest.fixture(scope="module")
def z():
    return np.load(GOLDEN)


def mats(z):
    X = sp.csr_matrix((z["X__data"], z["X__indices"], z["X__indptr"]), shape=tuple(z["X__shape"]))
    n = X.shape[0]
    H = sp.csr_matrix((z["H__data"], z["H__indices"], z["H__indptr"]), shape=(n, n))
    return X, H


def test_oracle_training_loop_reproduces_the_reference_fit(z):
    X, H = mats(z)
    Y = z["Y"]
    tr, dv, te = z["train_idx"], z["dev_idx"], z["test_idx"]
    net = go.GCNOracle(X, H, 2, False, tuple(z["regul_coefs"]))
    params = [z["init_%d" % i].copy() for i in range(4)]
    th, vh, best = go.fit_loop(net, params, tr, Y[tr].astype(np.int32), dv, Y[dv].astype(np.int32), 25,
                               early_stopping_max_down=1)
    ref_t, ref_v = z["f_train"], z["f_val"]
    assert len(th) == len(ref_t) == 21 and len(vh) == len(ref_v) == 4          # early stop at epoch 20 (:307)
    assert_close(np.array(th)[:, 0], ref_t[:, 0], atol=2e-6, rtol=2e-6, what="train loss per epoch")
    assert np.array_equal(np.array(th)[:, 1].astype(np.float32), ref_t[:, 1].astype(np.float32))   # accuracy per epoch
    assert_close(np.array(vh)[:, 0], ref_v[:, 0], atol=2e-6, rtol=2e-6, what="dev loss")
    assert np.allclose(np.array(vh)[:, 1], ref_v[:, 1])
    for i in range(4):
        assert_close(best[i], z["best_%d" % i], atol=1e-6, rtol=1e-4, what="best params %d" % i)
        assert_close(params[i], z["final_%d" % i], atol=1e-6, rtol=1e-4, what="restored params %d" % i)
        assert np.array_equal(z["best_%d" % i], z["final_%d" % i])              # :314 restores the snapshot
    assert_close(net.predict_proba(params, dv), z["proba_dev"], atol=1e-6, rtol=1e-4, what="predict_proba(dev)")
    assert np.array_equal(net.predict(params, te), z["predict_test"])


def test_reference_fit_semantics_read_off_the_golden_run(z):
    """facts of the reference run itself that the restatement relies on"""
    assert str(z["pickle_name"]) == "Xshape1_150_hidden_24_regul_0.0003_drop_0.5.pkl"      # mlpconv.py:310
    assert z["f_val"][0, 0] == z["f_val"][-1, 0]             # best = epoch 0 snapshot, re-validated at the end
    assert z["f_train"][0, 0] > z["f_train"][-1, 0]


@pytest.mark.gpu
def test_gpu_fit_reproduces_the_reference_fit(z, tmp_path):
    pytest.importorskip("torch")
    from graphconvgeo_b200.mlpconv import MLPCONV
    X, H = mats(z)
    Y = z["Y"]
    tr, dv, te = z["train_idx"], z["dev_idx"], z["test_idx"]
    clf = MLPCONV(n_epochs=25, batch_size=10, init_parameters=[z["init_%d" % i] for i in range(4)], complete_prob=False,
                  add_hidden=True, regul_coefs=list(z["regul_coefs"]), save_results=False, hidden_layer_size=int(z["hidden"][0]),
                  drop_out=False, dropout_coefs=[0.5, 0.5], early_stopping_max_down=1, loss_name='log',
                  nonlinearity='rectify', dtype='float32', model_dir=str(tmp_path), reorder=None)
    clf.fit(X, tr, dv, te, Y, H)
    assert clf._steps_done + (0 if clf._graph is None else 0) >= 1
    assert os.listdir(tmp_path) == [str(z["pickle_name"])]
    assert_close(clf.best_val[0], z["f_val"][-1, 0], atol=2e-6, rtol=2e-6, what="final dev loss")
    for i, p in enumerate(clf.get_param_values()):
        assert_close(p, z["final_%d" % i], atol=1e-6, rtol=1e-4, what="restored params %d" % i)
    assert_close(clf.predict_proba("dev"), z["proba_dev"], atol=1e-6, rtol=1e-4, what="predict_proba(dev)")
    assert np.array_equal(clf.predict("test"), z["predict_test"])
    assert abs(clf.accuracy("test", Y[te].astype("int32")) - float(z["accuracy_test"][0])) < 1e-7
