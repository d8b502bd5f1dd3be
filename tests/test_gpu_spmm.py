"""CUDA SpMM (gcg_spmm_csr_f32) against scipy's csr @ dense -- the routine Theano's S.dot
dispatches to (lasagne_layers.py:26,65,67,84).  Rows that are not split must be BIT-EXACT;
split (long) rows may differ by summation order only."""
import numpy as np
import pytest
import scipy.sparse as sp

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from util import assert_close, random_csr, to_dev  # noqa: E402


def run_spmm(A, B, thr=256, **kw):
    from graphconvgeo_b200 import ops
    from graphconvgeo_b200.sparse import CSRMatrix
    Ad = CSRMatrix.from_scipy(A, "cuda", long_row_threshold=thr)
    out = ops.spmm(Ad, to_dev(B), **kw)
    torch.cuda.synchronize()
    return out.cpu().numpy(), Ad


@pytest.mark.parametrize("F", [1, 3, 4, 16, 20, 64, 76, 100, 128, 256, 300, 600, 930, 1024])
@pytest.mark.parametrize("panel", [-2, -1, 0, 16, 64])
def test_bit_exact_vs_scipy(F, panel):
    rng = np.random.RandomState(F + panel + 2)
    A = random_csr(rng, 700, 500, 9)
    B = rng.standard_normal((500, F)).astype(np.float32)
    got, _ = run_spmm(A, B, thr=10 ** 6, panel_cols=panel)
    ref = np.asarray(A @ B, dtype=np.float32)
    assert np.array_equal(got, ref), "max diff %g" % np.abs(got - ref).max()


@pytest.mark.parametrize("panel", [-2, -1, 0, 16, 32, 128, 512])
def test_hub_rows_are_split_deterministically(panel):
    rng = np.random.RandomState(1)
    A = random_csr(rng, 400, 5000, 6, hub_rows=(0, 17, 399), hub_deg=3000)
    B = (rng.standard_normal((5000, 300)) * 0.1).astype(np.float32)
    got, Ad = run_spmm(A, B, thr=256, panel_cols=panel)
    info = Ad.plan_info()
    assert info["n_long_rows"] == 3 and info["n_segments"] == 3 * 12 and info["max_degree"] == 3000
    ref = np.asarray(A @ B, dtype=np.float32)
    short = np.ones(400, bool)
    short[[0, 17, 399]] = False
    assert np.array_equal(got[short], ref[short])
    assert_close(got[~short], ref[~short], atol=2e-5, what="hub rows")     # different summation order
    got2, _ = run_spmm(A, B, thr=256, panel_cols=panel)
    assert np.array_equal(got, got2), "split-row reduction must be deterministic"


def test_empty_matrix_and_empty_rows():
    rng = np.random.RandomState(2)
    A = sp.csr_matrix((50, 40), dtype=np.float32)
    B = rng.standard_normal((40, 8)).astype(np.float32)
    got, _ = run_spmm(A, B)
    assert np.array_equal(got, np.zeros((50, 8), np.float32))
    from graphconvgeo_b200 import ops
    bias = to_dev(rng.standard_normal(8))
    got, _ = run_spmm(A, B, bias=bias, act="tanh")
    assert_close(got, np.tile(np.tanh(bias.cpu().numpy()), (50, 1)))


@pytest.mark.parametrize("act", ["identity", "rectify", "tanh", "sigmoid"])
def test_fused_bias_activation(act):
    from oracle.gcn_oracle import ACTIVATIONS
    rng = np.random.RandomState(3)
    A = random_csr(rng, 300, 300, 7)
    B = (rng.standard_normal((300, 100)) * 0.3).astype(np.float32)
    b = rng.standard_normal(100).astype(np.float32)
    got, _ = run_spmm(A, B, bias=to_dev(b), act=act)
    pre = np.asarray(A @ B, dtype=np.float32) + b[None, :]
    ref = ACTIVATIONS[act](pre)
    if act in ("identity", "rectify"):
        assert np.array_equal(got, ref)
    else:
        assert_close(got, ref)


def test_fused_highway_gate_epilogue():
    """out = g*act(A.B+b) + (1-g)*H with the conv output optionally stored (north_star)."""
    rng = np.random.RandomState(4)
    n, F = 333, 300
    A = random_csr(rng, n, n, 8, hub_rows=(5,), hub_deg=300)
    B = (rng.standard_normal((n, F)) * 0.2).astype(np.float32)
    b = (rng.standard_normal(F) * 0.1).astype(np.float32)
    g = rng.rand(n, F).astype(np.float32)
    h = rng.standard_normal((n, F)).astype(np.float32)
    from graphconvgeo_b200 import ops
    conv = ops.alloc_mat(n, F, "cuda")
    got, _ = run_spmm(A, B, thr=64, bias=to_dev(b), act="rectify", gate=to_dev(g), carry=to_dev(h), conv_out=conv)
    hc = np.maximum(np.asarray(A @ B, dtype=np.float32) + b, 0)
    ref = g * hc + (np.float32(1) - g) * h
    assert_close(conv.cpu().numpy(), hc, atol=2e-6)
    assert_close(got, ref, atol=2e-6)
    got2, _ = run_spmm(A, B, thr=64, bias=to_dev(b), act="rectify", gate=to_dev(g), carry=to_dev(h))
    assert np.array_equal(got, got2)            # eval mode: same result, H' never written
    got3, _ = run_spmm(A, B, thr=64, bias=to_dev(b), act="rectify", gate=to_dev(g), carry=to_dev(h), panel_cols=-1)
    assert np.array_equal(got, got3)            # bulk-copy staged variant: same bits


STREAM_F = [4, 100, 128, 200, 256, 300, 384, 500, 600, 700, 1024]      # 1, 2, 3, 4, 5, 8 float4 per lane


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6, 7, 8])
def test_streaming_variants_are_bit_exact(variant):
    """Every kernel of gcg_spmm_stream.cu (shared-memory ring depths, register pipelines), short spans so that
    span / chunk / row boundaries fall everywhere, hub rows split into segments, empty rows, epilogues."""
    from graphconvgeo_b200 import _lib, ops
    from graphconvgeo_b200.sparse import CSRMatrix
    rng = np.random.RandomState(100 + variant)
    A = random_csr(rng, 900, 700, 11, hub_rows=(3, 450, 899), hub_deg=650, empty_frac=0.2)
    Ad = CSRMatrix.from_scipy(A, "cuda", long_row_threshold=64)
    L = _lib.lib()
    try:
        for span in (32, 100, 384):
            L.gcg_spmm_stream_tuning(variant, span, -1)
            for F in STREAM_F:
                B = rng.standard_normal((700, F)).astype(np.float32)
                b = rng.standard_normal(F).astype(np.float32)
                got = ops.spmm(Ad, to_dev(B), bias=to_dev(b), act="rectify", panel_cols=-2).cpu().numpy()
                want = ops.spmm(Ad, to_dev(B), bias=to_dev(b), act="rectify", panel_cols=0).cpu().numpy()
                assert np.array_equal(got, want), (variant, span, F)     # same bits as the register-gather kernel
                ref = np.maximum(np.asarray(A @ B, dtype=np.float32) + b[None, :], 0)
                short = np.diff(A.indptr) <= 64
                assert np.array_equal(got[short], ref[short]), (variant, span, F)   # and as scipy, unsplit rows
    finally:
        L.gcg_spmm_stream_tuning(0, 0, -1)


@pytest.mark.parametrize("sched", [([0, 1000], [3]), ([0, 1000], [-5]), ([0, 10, 10, 400, 1000], [1, 4, -2, 6]),
                                   ([0, 999, 1000], [64, 2])])
def test_streaming_block_panel_schedules_are_bit_exact(sched):
    from graphconvgeo_b200 import _lib, ops
    from graphconvgeo_b200.sparse import CSRMatrix
    rng = np.random.RandomState(7)
    A = random_csr(rng, 1000, 800, 9, hub_rows=(0, 500), hub_deg=700)
    Ad = CSRMatrix.from_scipy(A, "cuda", long_row_threshold=128)
    _lib.lib().gcg_spmm_stream_tuning(0, 64, -1)
    try:
        for F in (24, 300, 600, 930):
            B = to_dev(rng.standard_normal((800, F)).astype(np.float32))
            want = ops.spmm(Ad, B, panel_cols=0)
            Ad.set_schedule(None)
            assert torch.equal(ops.spmm(Ad, B, panel_cols=-2), want)
            Ad.set_schedule(*sched)
            assert torch.equal(ops.spmm(Ad, B, panel_cols=-2), want), (sched, F)
    finally:
        _lib.lib().gcg_spmm_stream_tuning(0, 0, -1)
    with pytest.raises(_lib.GcgError):
        Ad.set_schedule([0, 500], [1])                       # does not cover all rows


def test_streaming_gate_accumulate_and_row_slices():
    from graphconvgeo_b200 import ops
    from graphconvgeo_b200.sparse import CSRMatrix
    rng = np.random.RandomState(9)
    n, F = 333, 300
    A = random_csr(rng, n, n, 8, hub_rows=(5,), hub_deg=300)
    Ad = CSRMatrix.from_scipy(A, "cuda", long_row_threshold=64)
    B = to_dev((rng.standard_normal((n, F)) * 0.2).astype(np.float32))
    b = to_dev((rng.standard_normal(F) * 0.1).astype(np.float32))
    g, h = to_dev(rng.rand(n, F).astype(np.float32)), to_dev(rng.standard_normal((n, F)).astype(np.float32))
    c0, c2 = ops.alloc_mat(n, F, "cuda"), ops.alloc_mat(n, F, "cuda")
    want = ops.spmm(Ad, B, bias=b, act="tanh", gate=g, carry=h, conv_out=c0, panel_cols=0)
    got = ops.spmm(Ad, B, bias=b, act="tanh", gate=g, carry=h, conv_out=c2, panel_cols=-2)
    assert torch.equal(got, want) and torch.equal(c0, c2)
    acc0 = ops.spmm(Ad, B, panel_cols=0)
    acc2 = acc0.clone()
    ops.spmm(Ad, B, out=acc0, accumulate=True, act="rectify", panel_cols=0)
    ops.spmm(Ad, B, out=acc2, accumulate=True, act="rectify", panel_cols=-2)
    assert torch.equal(acc0, acc2)


def test_accumulate_mode():
    rng = np.random.RandomState(5)
    A1, A2 = random_csr(rng, 200, 150, 5), random_csr(rng, 200, 150, 5)
    B = rng.standard_normal((150, 64)).astype(np.float32)
    from graphconvgeo_b200 import ops
    from graphconvgeo_b200.sparse import CSRMatrix
    Bd = to_dev(B)
    out = ops.spmm(CSRMatrix.from_scipy(A1), Bd)
    ops.spmm(CSRMatrix.from_scipy(A2), Bd, out=out, accumulate=True, act="rectify")
    ref = np.maximum(np.asarray(A1 @ B, np.float32) + np.asarray(A2 @ B, np.float32), 0)
    assert_close(out.cpu().numpy(), ref)


def test_unaligned_operands_take_the_scalar_path():
    rng = np.random.RandomState(6)
    A = random_csr(rng, 120, 90, 6)
    Bfull = torch.from_numpy(rng.standard_normal((90, 37)).astype(np.float32)).cuda()     # ld 37: unaligned
    from graphconvgeo_b200 import ops
    from graphconvgeo_b200.sparse import CSRMatrix
    out = torch.empty(120, 37, device="cuda")
    ops.spmm(CSRMatrix.from_scipy(A), Bfull, out=out, act="rectify")
    ref = np.maximum(np.asarray(A @ Bfull.cpu().numpy(), np.float32), 0)
    assert np.array_equal(out.cpu().numpy(), ref)


def test_transpose_and_row_gather_plans():
    """X^T (Dot.grad) and A_hat[idx,:] (the fused row gather of lasagne_layers.py:88)."""
    rng = np.random.RandomState(7)
    from graphconvgeo_b200 import ops
    from graphconvgeo_b200.sparse import CSRMatrix
    X = random_csr(rng, 300, 180, 10)
    Xd = CSRMatrix.from_scipy(X)
    D = rng.standard_normal((300, 48)).astype(np.float32)
    got = ops.spmm(Xd.T, to_dev(D)).cpu().numpy()
    ref = np.asarray(sp.csr_matrix(X.T) @ D, dtype=np.float32)
    assert np.array_equal(got, ref)
    idx = rng.choice(300, size=77, replace=True).astype(np.int32)
    B = rng.standard_normal((180, 48)).astype(np.float32)
    got = ops.spmm(Xd.gather_rows(idx), to_dev(B)).cpu().numpy()
    assert np.array_equal(got, np.asarray(X @ B, np.float32)[idx])


def test_error_paths():
    from graphconvgeo_b200 import ops, _lib
    from graphconvgeo_b200.sparse import CSRMatrix
    rng = np.random.RandomState(8)
    A = CSRMatrix.from_scipy(random_csr(rng, 20, 30, 3))
    with pytest.raises(ValueError):
        ops.spmm(A, torch.zeros(31, 4, device="cuda"))
    with pytest.raises(TypeError):
        ops.spmm(A, torch.zeros(30, 4))                 # CPU tensor: no fallback
    B = torch.zeros(30, 4, device="cuda")
    with pytest.raises(_lib.GcgError, match="alias"):
        ops.spmm(CSRMatrix.from_scipy(random_csr(rng, 30, 30, 3)), B, out=B)


def test_property_random_shapes():
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=25, deadline=None)
    @given(st.integers(1, 300), st.integers(1, 300), st.integers(1, 160), st.integers(0, 12),
           st.sampled_from([-2, -1, 0, 16, 32, 64]), st.integers(0, 2 ** 31 - 1))
    def prop(n, k, F, deg, panel, seed):
        rng = np.random.RandomState(seed)
        A = random_csr(rng, n, k, deg, hub_rows=(0,), hub_deg=min(k, 70))
        B = rng.standard_normal((k, F)).astype(np.float32)
        got, _ = run_spmm(A, B, thr=32, panel_cols=panel)
        ref = np.asarray(A @ B, dtype=np.float32)
        assert_close(got, ref, atol=1e-5)
        assert np.array_equal(got[1:][np.diff(A.indptr)[1:] <= 32], ref[1:][np.diff(A.indptr)[1:] <= 32])
    prop()


def test_full_size_properties_twitter_world_shape():
    """BASELINE config 4 shape (1.4M nodes): size-independent properties instead of an oracle run:
    A_hat.1 = row sums, linearity, and x^T(A y) = y^T(A x) (A_hat symmetric)."""
    from graphconvgeo_b200 import ops
    from graphconvgeo_b200.sparse import CSRMatrix, build_ahat_host
    from graphconvgeo_b200.synth import powerlaw_graph
    n, F = 1_400_000, 64
    A = build_ahat_host(powerlaw_graph(n, 10, seed=3))
    Ad = CSRMatrix.from_scipy(A)
    info = Ad.plan_info()
    assert info["n_long_rows"] > 0
    ones = torch.ones(n, F, device="cuda")
    rs = ops.spmm(Ad, ones)
    ref = np.add.reduceat(A.data.astype(np.float64), A.indptr[:-1])
    assert_close(rs[:, 0].cpu().numpy(), ref.astype(np.float32), atol=1e-5, what="row sums")
    assert torch.equal(rs[:, 0], rs[:, F - 1])
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(n, F, device="cuda", generator=g)
    y = torch.randn(n, F, device="cuda", generator=g)
    ax, ay = ops.spmm(Ad, x), ops.spmm(Ad, y)
    axy = ops.spmm(Ad, 2.0 * x - 0.5 * y)
    lin = (2.0 * ax - 0.5 * ay - axy).abs().max().item()
    assert lin < 5e-5, lin
    s1 = (y.double() * ax.double()).sum().item()
    s2 = (x.double() * ay.double()).sum().item()
    assert abs(s1 - s2) <= 1e-6 * max(abs(s1), 1.0), (s1, s2)
    # panel-major execution and the bulk-copy staged kernel give the same bits as whole-row execution
    assert torch.equal(ops.spmm(Ad, x, panel_cols=16), ax)
    assert torch.equal(ops.spmm(Ad, x, panel_cols=-1), ops.spmm(Ad, x, panel_cols=0))
    # so does the nnz-balanced streaming kernel, whole rows and under a block x panel schedule
    assert torch.equal(ops.spmm(Ad, x, panel_cols=-2), ax)
    Ad.set_schedule([0, 300_000, 300_000, 1_000_001, n], [2, 1, -2, 1])
    assert torch.equal(ops.spmm(Ad, x, panel_cols=-2), ax)


def test_document_blocked_transpose_product_matches_plain():
    """sparse.BlockedRows: heavy rows of X^T processed per column block with accumulation == one plain SpMM."""
    from graphconvgeo_b200 import ops
    from graphconvgeo_b200.sparse import BlockedRows, CSRMatrix
    rng = np.random.RandomState(11)
    light = random_csr(rng, 400, 6000, 3)
    heavy = sp.random(12, 6000, density=0.4, format="csr", dtype=np.float32, random_state=rng)
    XT = sp.vstack([light[:200], heavy, light[200:]]).tocsr()
    XT.sort_indices()
    D = (rng.standard_normal((6000, 96)) * 0.1).astype(np.float32)
    Xd = CSRMatrix.from_scipy(XT)
    br = BlockedRows(Xd, F=96, block_mb=0)          # block_cols = 1024 -> 6 column blocks
    assert br.n_blocks == 6 and br.n_heavy >= 12 and len(br.blocks) == 6
    out = ops.alloc_mat(XT.shape[0], 96, "cuda")
    out.fill_(123.0)
    br.product(to_dev(D), out)
    ref = np.asarray(XT @ D, dtype=np.float32)
    got = out.cpu().numpy()
    lightrows = np.ones(XT.shape[0], bool)
    lightrows[br.heavy_ids] = False
    assert np.array_equal(got[lightrows], ref[lightrows])          # light rows: same kernel, same order
    assert_close(got[~lightrows], ref[~lightrows], atol=2e-5)      # heavy rows: block-wise summation order
    assert np.array_equal(got, br.product(to_dev(D), ops.alloc_mat(XT.shape[0], 96, "cuda")).cpu().numpy())


def test_stacked_document_blocks_equal_the_block_loop(monkeypatch):
    """BlockedRows: one streaming launch over all (block, heavy row) pieces + slab sum == the per-block loop."""
    from graphconvgeo_b200 import ops
    from graphconvgeo_b200.sparse import BlockedRows, CSRMatrix
    rng = np.random.RandomState(21)
    light = random_csr(rng, 300, 5000, 3)
    heavy = sp.random(20, 5000, density=0.5, format="csr", dtype=np.float32, random_state=rng)
    XT = sp.vstack([light[:100], heavy, light[100:]]).tocsr()
    XT.sort_indices()
    D = to_dev((rng.standard_normal((5000, 128)) * 0.1).astype(np.float32))
    Xd = CSRMatrix.from_scipy(XT, long_row_threshold=512)
    monkeypatch.setenv("GCG_XT_STACKED", "1")
    a = BlockedRows(Xd, F=128, block_mb=0)
    assert a.stacked is not None and a.stacked.shape[0] == a.n_heavy * len(a.blocks)
    monkeypatch.setenv("GCG_XT_STACKED", "0")
    b = BlockedRows(Xd, F=128, block_mb=0)
    assert b.stacked is None
    ra = a.product(D, ops.alloc_mat(XT.shape[0], 128, "cuda")).cpu().numpy()
    rb = b.product(D, ops.alloc_mat(XT.shape[0], 128, "cuda")).cpu().numpy()
    ref = np.asarray(XT @ D.cpu().numpy(), dtype=np.float32)
    assert_close(ra, ref, atol=2e-5)
    assert_close(ra, rb, atol=2e-5)
    for kern in ("gather", "stream"):
        monkeypatch.setenv("GCG_XT_STACKED", "1")
        monkeypatch.setenv("GCG_XT_STACKED_KERNEL", kern)
        c = BlockedRows(Xd, F=128, block_mb=0)
        assert np.array_equal(c.product(D, ops.alloc_mat(XT.shape[0], 128, "cuda")).cpu().numpy(), ra)
