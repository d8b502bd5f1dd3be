"""End-to-end parity of the GCN (MLPCONV, mlpconv.py:121-346) with the oracle: per-layer
activations and gradients within 1e-6 + 1e-4*|ref| (north_star), identical argmax except
exact ties, a multi-step Adam trajectory, CUDA-graph replay == eager, and fit()."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import gcn_oracle as go  # noqa: E402
from util import assert_close  # noqa: E402


def workload(name="tiny", **kw):
    from graphconvgeo_b200 import synth
    return synth.make_workload(name, **kw)


def make_model(w, n_layers, highway, params, idx, cuda_graph=False, act="rectify", reg=(1e-4, 2e-4), reorder=None):
    from graphconvgeo_b200.mlpconv import MLPCONV
    m = MLPCONV(n_epochs=1, regul_coefs=list(reg), hidden_layer_size=w.hidden, drop_out=False,
                n_layers=n_layers, highway=highway, init_parameters=[p.copy() for p in params],
                cuda_graph=cuda_graph, nonlinearity=act, reorder=reorder)
    m.prepare(w.X, idx, w.dev_indices, w.test_indices, w.Y, w.A_hat)
    return m


CASES = [(2, False, "rectify", False), (2, False, "rectify", True), (3, True, "rectify", False),
         (4, True, "tanh", True), (3, False, "rectify", False)]


@pytest.mark.parametrize("n_layers,highway,act,dup", CASES)
def test_one_training_step_matches_oracle(n_layers, highway, act, dup):
    w = workload()
    rng = np.random.RandomState(5)
    idx = w.train_indices
    if dup:
        idx = rng.choice(w.train_indices, size=len(w.train_indices)).astype(np.int32)     # tensormain.py:226
    params = go.init_params(rng, w.X.shape[1], w.hidden, w.n_classes, n_layers, highway)
    for p in params:
        if p.ndim == 1:
            p[...] = (rng.standard_normal(p.shape) * 0.05).astype(np.float32)
    reg = (1e-4, 2e-4)
    net = go.GCNOracle(w.X, w.A_hat, n_layers, highway, reg, act)
    y = w.Y[idx].astype(np.int32)
    loss, acc, grads, cache = net.loss_and_grads(params, idx, y)
    m = make_model(w, n_layers, highway, params, idx, act=act, reg=reg)
    m.f_train()
    torch.cuda.synchronize()
    l_gpu, a_gpu = m.train_results()
    assert abs(l_gpu - float(loss)) <= 1e-6 + 1e-5 * abs(float(loss))
    assert abs(a_gpu - acc) < 1e-6
    # per-layer activations
    convs = [ly for ly in m.layers]
    for i, ly in enumerate(convs[:-1]):
        assert_close(m.node_rows(ly._out), cache["A"][i], what="activation of layer %d" % (i + 1))
    assert_close(m.l_out._out.cpu().numpy(), cache["logits"], atol=2e-6, what="logits")
    # per-parameter gradients (the reg sub-gradient is folded into the GPU Adam kernel, so add it here)
    gpu_grads = m.get_grad_values()
    k = 0
    for li, ly in enumerate(m.layers):
        for name, t, tags in ly.params:
            g = gpu_grads[k]
            if tags.get("regularizable"):
                c = np.float32(reg[0] if ly is m.l_out else reg[1])
                g = g + np.float32(0.5) * c * (np.sign(params[k]) + np.float32(2) * params[k])
            assert_close(g, grads[k], what="grad of layer %d %s" % (li + 1, name))
            k += 1
    # parameters after the Adam step
    st = go.AdamState(params)
    go.adam_step(params, grads, st)
    for p_gpu, p in zip(m.get_param_values(), params):
        assert_close(p_gpu, p, atol=2e-6, what="params after step")


@pytest.mark.parametrize("n_layers,highway", [(2, False), (3, True)])
def test_five_step_trajectory_and_predictions(n_layers, highway):
    w = workload()
    rng = np.random.RandomState(11)
    params = go.init_params(rng, w.X.shape[1], w.hidden, w.n_classes, n_layers, highway)
    idx = w.train_indices
    y = w.Y[idx].astype(np.int32)
    net = go.GCNOracle(w.X, w.A_hat, n_layers, highway, (1e-4, 2e-4))
    ref_params = [p.copy() for p in params]
    hist = go.train_epochs(net, ref_params, idx, y, 5)
    m = make_model(w, n_layers, highway, params, idx, cuda_graph=True)     # epochs 2..5 replay the graph
    for step in range(5):
        m.f_train()
        l, a = m.train_results()
        assert abs(l - hist[step][0]) <= 2e-5 * abs(hist[step][0]), (step, l, hist[step])
        assert abs(a - hist[step][1]) <= 2.0 / len(idx)
    assert m._graph is not None
    for p_gpu, p in zip(m.get_param_values(), ref_params):
        assert_close(p_gpu, p, atol=1e-5, rtol=1e-3, what="params after 5 steps")
    # predictions with identical parameters: argmax identical except exact logit ties
    from graphconvgeo_b200 import lasagne_layers as L
    L.set_all_param_values(m.l_out, ref_params)
    for part, pidx in (("dev", w.dev_indices), ("test", w.test_indices), ("train", idx)):
        proba = m.predict_proba(part)
        ref = net.predict_proba(ref_params, pidx)
        assert_close(proba, ref, atol=1e-6, what="predict_proba " + part)
        pred = m.predict(part)
        assert pred.dtype == np.int64
        top2 = np.sort(ref, axis=1)[:, -2:]
        tie = (top2[:, 1] - top2[:, 0]) <= 1e-6
        assert np.array_equal(pred[~tie], ref.argmax(-1)[~tie])
    acc = m.accuracy("test", w.Y[w.test_indices].astype(np.int32))
    _, ref_acc = net.loss_acc(ref_params, w.test_indices, w.Y[w.test_indices].astype(np.int32))
    assert abs(acc - ref_acc) <= 2.0 / len(w.test_indices)


def test_cuda_graph_replay_equals_eager():
    w = workload()
    rng = np.random.RandomState(3)
    params = go.init_params(rng, w.X.shape[1], w.hidden, w.n_classes, 3, True)
    a = make_model(w, 3, True, params, w.train_indices, cuda_graph=True)
    b = make_model(w, 3, True, params, w.train_indices, cuda_graph=False)
    a.native_epoch = b.native_epoch = False     # the layer code itself, captured vs eager (the native epoch: next test)
    for _ in range(4):
        a.f_train()
        b.f_train()
    torch.cuda.synchronize()
    for pa, pb in zip(a.get_param_values(), b.get_param_values()):
        assert np.array_equal(pa, pb)            # same kernels, same order -> same bits
    assert a.train_results() == b.train_results()


@pytest.mark.parametrize("n_layers,highway,graph", [(2, False, False), (3, True, False), (3, True, True), (4, True, False)])
def test_native_epoch_program_equals_the_layer_code(n_layers, highway, graph):
    """the epoch as one native object (gcg_epoch_*, include/gcg.h) -- every libgcg call of
    f_train recorded once, replayed from C++ with one call per epoch (optionally inside a CUDA graph) -- gives the
    same parameters, loss and accuracy, bit for bit, as running the layer code call by call from Python."""
    from graphconvgeo_b200 import ops
    w = workload()
    rng = np.random.RandomState(5)
    params = go.init_params(rng, w.X.shape[1], w.hidden, w.n_classes, n_layers, highway)
    a = make_model(w, n_layers, highway, params, w.train_indices, cuda_graph=graph)
    a.native_epoch = True
    b = make_model(w, n_layers, highway, params, w.train_indices, cuda_graph=False)
    b.native_epoch = False
    for step in range(5):
        a.f_train()
        b.f_train()
        torch.cuda.synchronize()
        assert a.train_results() == b.train_results(), step
    assert a._program is not None and (a._graph is not None) == graph and b._program is None
    for pa, pb in zip(a.get_param_values(), b.get_param_values()):
        assert np.array_equal(pa, pb)
    names = a._program.call_names()
    assert len(names) == len(a._program) > 10
    assert names[-1] == "gcg_adam_step_f32" and "gcg_softmax_ce_f32" in names and "gcg_spmm_csr_f32" in names
    assert sum(n == "gcg_spmm_csr_f32" for n in names) >= 2 * n_layers        # X.W1, A_hat products, their transposes
    # one C call enqueues them all: the library's launch counter advances by the same amount as an eager epoch
    ops.launch_count(reset=True)
    b.f_train()
    eager = ops.launch_count(reset=True)
    if not graph:
        a.f_train()
        assert ops.launch_count(reset=True) == eager
    # evaluation between epochs does not disturb the recorded program
    va = a.f_val(a.y_dev_dev, a.ti["dev"])
    vb = b.f_val(b.y_dev_dev, b.ti["dev"])
    if graph:
        a.f_train()          # (b took one more step above, for the launch count)
    a.f_train()
    b.f_train()
    torch.cuda.synchronize()
    assert a.train_results() == b.train_results()
    assert va[1] == pytest.approx(vb[1])


def test_epoch_program_api_errors():
    from graphconvgeo_b200 import _lib, ops
    L = _lib.lib()
    p = ops.EpochProgram()
    assert len(p) == 0 and p.call_names() == []
    p.run()                                                     # an empty program is a no-op
    with p.record():
        q = ops.EpochProgram()
        with pytest.raises(_lib.GcgError, match="already recording"):
            with q.record():
                pass
        with pytest.raises(_lib.GcgError, match="still being recorded"):
            p.run()
        x = torch.arange(12, dtype=torch.float32, device="cuda")
        out = torch.zeros(1, dtype=torch.float32, device="cuda")
        ops.sum_scaled(x, 0.5, out=out)
    assert p.call_names() == ["gcg_sum_f32"] and float(out.item()) == 33.0
    x.mul_(2.0)
    p.run()
    assert float(out.item()) == 66.0
    with pytest.raises(_lib.GcgError):
        _lib.check(L.gcg_epoch_run(None, None), "gcg_epoch_run")


def test_fit_learns_and_keeps_the_reference_interface(tmp_path):
    from graphconvgeo_b200.mlpconv import MLPCONV
    w = workload()
    clf = MLPCONV(n_epochs=61, batch_size=500, init_parameters=None, complete_prob=False, add_hidden=True,
                  regul_coefs=[1e-6, 1e-6], save_results=False, hidden_layer_size=w.hidden, drop_out=False,
                  dropout_coefs=[0.5, 0.5], early_stopping_max_down=5, loss_name='log', nonlinearity='rectify',
                  dtype='float32', seed=1, model_dir=str(tmp_path))
    clf.fit(w.X, w.train_indices, w.dev_indices, w.test_indices, w.Y, w.A_hat)       # tensormain.py:237
    acc = clf.accuracy(dataset_partition='test', y_true=w.Y[w.test_indices].astype('int32'))
    y_pred = clf.predict(dataset_partition='test')
    assert y_pred.shape == (len(w.test_indices),)
    assert acc > 3.0 / w.n_classes, acc          # far above chance: labels follow the graph communities
    assert abs(acc - np.mean(y_pred == w.Y[w.test_indices])) < 1e-6
    files = list(tmp_path.iterdir())
    assert len(files) == 1 and files[0].name.startswith("Xshape1_%d_hidden_%d" % (w.X.shape[1], w.hidden))
    import pickle
    best = pickle.load(open(files[0], "rb"))
    assert [b.shape for b in best] == [p.shape for p in clf.get_param_values()]
    with pytest.raises(ValueError):
        clf.predict("validation")


def test_geotext_shape_reference_network_step():
    """BASELINE config 1/2 shape (GEOTEXT: 9,475 users, 9k vocab, 128 regions, hid 300)."""
    w = workload("geotext")
    rng = np.random.RandomState(0)
    for n_layers, highway in ((2, False), (3, True)):
        params = go.init_params(rng, w.X.shape[1], w.hidden, w.n_classes, n_layers, highway)
        net = go.GCNOracle(w.X, w.A_hat, n_layers, highway, (1e-6, 1e-6))
        y = w.Y[w.train_indices].astype(np.int32)
        loss, acc, grads, cache = net.loss_and_grads(params, w.train_indices, y)
        m = make_model(w, n_layers, highway, params, w.train_indices, reg=(1e-6, 1e-6),
                       reorder="labels" if highway else None)
        m.f_train()
        l_gpu, a_gpu = m.train_results()
        assert abs(l_gpu - float(loss)) <= 1e-5 * abs(float(loss))
        flips = 0
        for i, ly in enumerate(m.layers[:-1]):
            assert_close(m.node_rows(ly._out), cache["A"][i], what="activation %d" % i)
            pre_gpu = m.node_rows(ly._Hc) if getattr(ly, "_Hc", None) is not None else m.node_rows(ly._out)
            pre_ref = cache["Hc"][i] if cache["Hc"][i] is not None else cache["A"][i]
            flips += int(np.count_nonzero((pre_gpu > 0) != (pre_ref > 0)))
        # Node reordering permutes the columns of A_hat, hence the CSR summation order inside a row: a
        # pre-activation within an ulp of zero can land on the other side of the relu kink, and relu' then differs
        # for that ONE element (both evaluations are valid float32).  Each such flip moves a gradient entry by at
        # most |dO| of that element (~1e-6 here), so the bound is widened by that much per flip -- only then.
        atol = 1e-6 + 2e-6 * flips
        gpu_grads = m.get_grad_values()
        k = 0
        for ly in m.layers:
            for name, t, tags in ly.params:
                g = gpu_grads[k]
                if tags.get("regularizable"):
                    g = g + np.float32(0.5e-6) * (np.sign(params[k]) + np.float32(2) * params[k])
                assert_close(g, grads[k], atol=atol, what="grad %d (%d relu flips)" % (k, flips))
                k += 1


@pytest.mark.parametrize("mode", ["labels", "degree", "random"])
def test_node_reordering_changes_no_result(mode):
    """The GCN is permutation equivariant: running on (P A P^T, P X) with mapped target indices gives
    the same loss, gradients, activations (in original order) and predictions."""
    w = workload()
    rng = np.random.RandomState(21)
    params = go.init_params(rng, w.X.shape[1], w.hidden, w.n_classes, 3, True)
    idx = rng.choice(w.train_indices, size=500).astype(np.int32)
    y = w.Y[idx].astype(np.int32)
    net = go.GCNOracle(w.X, w.A_hat, 3, True, (1e-4, 2e-4))
    loss, acc, grads, cache = net.loss_and_grads(params, idx, y)
    reorder = rng.permutation(w.X.shape[0]).astype(np.int32) if mode == "random" else mode
    m = make_model(w, 3, True, params, idx, reorder=reorder)
    assert m.node_order is not None
    m.f_train()
    l_gpu, a_gpu = m.train_results()
    assert abs(l_gpu - float(loss)) <= 1e-5 * abs(float(loss)) and abs(a_gpu - acc) < 1e-6
    for i, ly in enumerate(m.layers[:-1]):
        assert_close(m.node_rows(ly._out), cache["A"][i], what="activation %d" % i)
    assert_close(m.l_out._out.cpu().numpy(), cache["logits"], atol=2e-6)      # idx order is untouched
    st = go.AdamState(params)
    go.adam_step(params, grads, st)
    for p_gpu, p in zip(m.get_param_values(), params):
        assert_close(p_gpu, p, atol=2e-6)
    ref = net.predict_proba(params, w.test_indices)
    assert_close(m.predict_proba("test"), ref, atol=1e-6)
    emb = m.get_embedding(w.dev_indices[:7])
    assert emb.shape == (7, w.hidden)


@pytest.mark.parametrize("n_layers,highway", [(2, False), (3, True)])
def test_training_step_with_tensor_core_gemms(n_layers, highway):
    """Same parity check with every dense projection forced onto the tcgen05 3xTF32 engine (the engine
    the large benchmarks use).  The tensor core's round-toward-zero accumulation costs ~3e-5 of the
    output magnitude, so the absolute part of the tolerance is 4e-6 here instead of 1e-6."""
    from graphconvgeo_b200 import _lib, ops
    if not (_lib.lib().gcg_gemm_tc_available()):
        pytest.skip("tcgen05 engine unavailable")
    w = workload("geotext")
    rng = np.random.RandomState(2)
    params = go.init_params(rng, w.X.shape[1], w.hidden, w.n_classes, n_layers, highway)
    net = go.GCNOracle(w.X, w.A_hat, n_layers, highway, (1e-6, 1e-6))
    y = w.Y[w.train_indices].astype(np.int32)
    loss, acc, grads, cache = net.loss_and_grads(params, w.train_indices, y)
    ops.set_gemm_mode("tf32x3")
    try:
        m = make_model(w, n_layers, highway, params, w.train_indices, reg=(1e-6, 1e-6))
        ops.launch_count(reset=True)
        m.f_train()
        torch.cuda.synchronize()
    finally:
        ops.set_gemm_mode("auto")
    l_gpu, a_gpu = m.train_results()
    assert abs(l_gpu - float(loss)) <= 1e-5 * abs(float(loss))
    for i, ly in enumerate(m.layers[:-1]):
        assert_close(m.node_rows(ly._out), cache["A"][i], atol=4e-6, what="activation %d" % i)
    assert_close(m.l_out._out.cpu().numpy(), cache["logits"], atol=4e-6, what="logits")
    gpu_grads = m.get_grad_values()
    k = 0
    for ly in m.layers:
        for name, t, tags in ly.params:
            g = gpu_grads[k]
            if tags.get("regularizable"):
                g = g + np.float32(0.5e-6) * (np.sign(params[k]) + np.float32(2) * params[k])
            assert_close(g, grads[k], atol=4e-6, what="grad %d (%s)" % (k, name))
            k += 1


def test_model_with_more_regions_than_hidden_units_uses_propagate_first():
    """Twitter-World shape in miniature: C > hidden, so the output layer propagates first."""
    w = workload(bucket=4, hidden=24)
    assert w.n_classes > w.hidden
    rng = np.random.RandomState(8)
    params = go.init_params(rng, w.X.shape[1], w.hidden, w.n_classes, 3, True)
    idx = rng.choice(w.train_indices, size=400).astype(np.int32)
    y = w.Y[idx].astype(np.int32)
    net = go.GCNOracle(w.X, w.A_hat, 3, True, (1e-4, 2e-4))
    loss, acc, grads, cache = net.loss_and_grads(params, idx, y)
    m = make_model(w, 3, True, params, idx)
    assert m.l_out.propagate_first
    m.f_train()
    l_gpu, a_gpu = m.train_results()
    assert abs(l_gpu - float(loss)) <= 1e-5 * abs(float(loss))
    assert_close(m.l_out._out.cpu().numpy(), cache["logits"], atol=2e-6, what="logits")
    st = go.AdamState(params)
    go.adam_step(params, grads, st)
    for p_gpu, p in zip(m.get_param_values(), params):
        assert_close(p_gpu, p, atol=2e-6, what="params after step")
