"""Epilogue, reduction, head and optimiser kernels against the NumPy oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from util import assert_close, to_dev  # noqa: E402


@pytest.mark.parametrize("n,F", [(1, 1), (7, 3), (1000, 300), (5000, 600), (12345, 129), (64, 1024)])
def test_colsum(n, F):
    from graphconvgeo_b200 import ops
    rng = np.random.RandomState(n + F)
    X = rng.standard_normal((n, F)).astype(np.float32)
    got = ops.colsum(to_dev(X)).cpu().numpy()
    assert_close(got, X.astype(np.float64).sum(0), atol=1e-5 * np.sqrt(n))
    # unaligned leading dimension -> scalar path, same answer within tolerance; deterministic
    Xu = torch.from_numpy(X).cuda()
    g2 = ops.colsum(Xu).cpu().numpy()
    assert_close(g2, X.astype(np.float64).sum(0), atol=1e-5 * np.sqrt(n))
    assert np.array_equal(g2, ops.colsum(Xu).cpu().numpy())


@pytest.mark.parametrize("act", ["rectify", "tanh", "sigmoid", "identity"])
def test_act_bwd(act):
    from graphconvgeo_b200 import ops
    from oracle.gcn_oracle import ACTIVATIONS
    rng = np.random.RandomState(0)
    P = rng.standard_normal((301, 77)).astype(np.float32)
    A = ACTIVATIONS[act](P).astype(np.float32)
    dA = rng.standard_normal(P.shape).astype(np.float32)
    got = ops.act_bwd(to_dev(dA), to_dev(A), act).cpu().numpy()
    d = {"rectify": (P > 0).astype(np.float32), "tanh": 1 - A * A, "sigmoid": A * (1 - A), "identity": np.ones_like(P)}[act]
    assert_close(got, dA * d)


def test_highway_bwd():
    from graphconvgeo_b200 import ops
    rng = np.random.RandomState(1)
    n, F = 257, 300
    dO = rng.standard_normal((n, F)).astype(np.float32)
    g = rng.rand(n, F).astype(np.float32)
    Hc = np.maximum(rng.standard_normal((n, F)), 0).astype(np.float32)
    Hin = rng.standard_normal((n, F)).astype(np.float32)
    dP, dG, dH = ops.highway_bwd(to_dev(dO), to_dev(g), to_dev(Hc), to_dev(Hin), "rectify")
    assert_close(dP.cpu().numpy(), g * dO * (Hc > 0))
    assert_close(dG.cpu().numpy(), dO * (Hc - Hin) * g * (1 - g), atol=2e-6)
    assert_close(dH.cpu().numpy(), (1 - g) * dO)


def test_highway_mix_standalone_equals_fused_formula():
    from graphconvgeo_b200 import ops
    rng = np.random.RandomState(7)
    n, F = 123, 77
    hc, g, h = (rng.standard_normal((n, F)).astype(np.float32) for _ in range(3))
    g = np.abs(g) % 1
    got = ops.highway_mix(to_dev(hc), to_dev(g), to_dev(h)).cpu().numpy()
    assert np.array_equal(got, g * hc + (np.float32(1) - g) * h)


@pytest.mark.parametrize("n,C", [(1, 2), (33, 7), (500, 128), (300, 930), (64, 1024), (10, 2000)])
def test_softmax_ce_head(n, C):
    from graphconvgeo_b200 import ops
    rng = np.random.RandomState(n + C)
    Lg = (rng.standard_normal((n, C)) * 3).astype(np.float32)
    Lg[0, :] = 1.5                                   # exact ties: first maximum wins (np.argmax)
    if n > 2:
        Lg[2, C // 2] = Lg[2].max()                  # duplicate maximum
    y = rng.randint(0, C, n).astype(np.int32)
    d = "cuda"
    probs = ops.alloc_mat(n, C, d)
    G = ops.alloc_mat(n, C, d)
    ce = torch.empty(n, device=d)
    hit = torch.empty(n, device=d)
    pred = torch.empty(n, dtype=torch.int64, device=d)
    ops.softmax_ce(to_dev(Lg), y=torch.from_numpy(y).cuda(), probs=probs, grad=G, ce=ce, hit=hit, pred=pred, denom=n)
    m = Lg.max(1, keepdims=True)
    e = np.exp((Lg - m).astype(np.float64))
    p = e / e.sum(1, keepdims=True)
    assert_close(probs.cpu().numpy(), p, atol=1e-7)
    assert np.array_equal(pred.cpu().numpy(), Lg.argmax(1))
    assert np.array_equal(hit.cpu().numpy(), (Lg.argmax(1) == y).astype(np.float32))
    assert_close(ce.cpu().numpy(), -np.log(p[np.arange(n), y]), atol=2e-6)
    Gr = p.copy()
    Gr[np.arange(n), y] -= 1
    assert_close(G.cpu().numpy(), Gr / n, atol=1e-7)
    # predict-only call (no labels)
    pred2 = torch.empty(n, dtype=torch.int64, device=d)
    ops.softmax_ce(to_dev(Lg), pred=pred2)
    assert torch.equal(pred, pred2)


def test_sum_scatter_gather():
    from graphconvgeo_b200 import ops
    rng = np.random.RandomState(2)
    x = rng.standard_normal(100001).astype(np.float32)
    s = ops.sum_scaled(to_dev(x), 0.5).item()
    assert abs(s - 0.5 * x.astype(np.float64).sum()) < 1e-2
    N, C, n_idx = 500, 37, 800
    idx = rng.randint(0, N, n_idx).astype(np.int32)       # heavy duplicates (tensormain.py:226)
    G = rng.standard_normal((n_idx, C)).astype(np.float32)
    ptr, pos = ops.scatter_positions(idx, N, "cuda")
    dP = ops.scatter_rows(to_dev(G), ptr, pos, N).cpu().numpy()
    ref = np.zeros((N, C), np.float64)
    np.add.at(ref, idx, G.astype(np.float64))
    assert_close(dP, ref, atol=2e-6)
    assert np.array_equal(dP, ops.scatter_rows(to_dev(G), ptr, pos, N).cpu().numpy())
    X = rng.standard_normal((N, C)).astype(np.float32)
    got = ops.gather_rows(to_dev(X), torch.from_numpy(idx).cuda()).cpu().numpy()
    assert np.array_equal(got, X[idx])
    # put_rows: the inverse placement for distinct indices (aligned and unaligned widths); other rows untouched
    for Cc in (37, 600):
        sel = rng.permutation(N)[:123].astype(np.int32)
        src = rng.standard_normal((len(sel), Cc)).astype(np.float32)
        base = rng.standard_normal((N, Cc)).astype(np.float32)
        dst = to_dev(base)
        ops.put_rows(to_dev(src), torch.from_numpy(sel).cuda(), dst)
        want = base.copy()
        want[sel] = src
        assert np.array_equal(dst.cpu().numpy(), want)
    with pytest.raises(ValueError):
        ops.put_rows(to_dev(src), torch.from_numpy(sel[:5]).cuda(), dst)


def test_adam_matches_lasagne_adam_oracle():
    from graphconvgeo_b200 import ops
    from oracle import gcn_oracle as go
    rng = np.random.RandomState(3)
    shapes = [(300, 64), (64,), (64, 17), (17,), (5000, 3)]
    reg = [1e-3, 0.0, 2e-3, 0.0, 5e-4]
    P = [rng.standard_normal(s).astype(np.float32) * 0.1 for s in shapes]
    P[0][0, :5] = 0.0                                    # sign(0) = 0 in the L1 sub-gradient
    pd = [torch.from_numpy(p.copy()).cuda() for p in P]
    gd = [torch.zeros_like(p) for p in pd]
    adam = ops.Adam(pd, gd, reg)
    en = ops.ElasticNet(pd, reg)
    st = go.AdamState(P)
    for step in range(4):
        G = [rng.standard_normal(s).astype(np.float32) * 0.01 for s in shapes]
        for g_d, g in zip(gd, G):
            g_d.copy_(torch.from_numpy(g))
        reg_ref = sum(0.5 * c * (np.abs(p).sum(dtype=np.float64) + (p.astype(np.float64) ** 2).sum())
                      for p, c in zip(P, reg))
        assert abs(en().item() - reg_ref) <= 1e-5 * max(reg_ref, 1)
        full = [g + np.float32(0.5 * c) * (np.sign(p) + np.float32(2) * p) for g, p, c in zip(G, P, reg)]
        go.adam_step(P, full, st)
        r = adam.step()
        assert abs(r.item() - reg_ref) <= 1e-5 * max(reg_ref, 1)
        for p_d, p in zip(pd, P):
            assert_close(p_d.cpu().numpy(), p, atol=1e-6, what="step %d" % step)
    assert adam.t[0].item() == 4.0


@pytest.mark.parametrize("n,F,P", [(5, 7, 2), (1000, 600, 8), (333, 1024, 8), (64, 30, 4)])
def test_pack_unpack_column_slices(n, F, P):
    from graphconvgeo_b200 import ops
    rng = np.random.RandomState(n + F)
    X = rng.standard_normal((n, F)).astype(np.float32)
    Fp = (-(-F // P) + 3) // 4 * 4
    dst = torch.full((P, n, Fp), 7.0, device="cuda")
    ops.pack_cols(to_dev(X), P, Fp, dst)
    ref = np.zeros((P, n, Fp), np.float32)
    for q in range(P):
        w = max(0, min(Fp, F - q * Fp))
        ref[q, :, :w] = X[:, q * Fp:q * Fp + w]
    assert np.array_equal(dst.cpu().numpy(), ref)
    back = ops.alloc_mat(n, F, "cuda")
    ops.unpack_cols(dst, P, Fp, back)
    assert np.array_equal(back.cpu().numpy(), X)
