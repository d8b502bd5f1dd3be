"""Dense projections (gcg_gemm_f32) against NumPy: T.dot of lasagne_layers.py:82 and the two
gradient products theano.grad derives from it."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from util import assert_close, to_dev  # noqa: E402

MODES = ["fma"]


def tc_modes():
    from graphconvgeo_b200 import _lib
    L = _lib.lib()
    if hasattr(L, "gcg_gemm_tc_available") and L.gcg_gemm_tc_available():
        return ["fma", "tf32x3"]
    return ["fma"]


def ref_gemm(A, B, ta, tb):
    a = A.T if ta else A
    b = B.T if tb else B
    return (a.astype(np.float64) @ b.astype(np.float64))


@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (128, 128, 16), (130, 70, 33), (257, 300, 600), (1000, 930, 300),
                                   (64, 1024, 77), (300, 128, 5000)])
def test_gemm_shapes(ta, tb, M, N, K):
    from graphconvgeo_b200 import ops
    rng = np.random.RandomState(M + N + K)
    s = 1.0 / np.sqrt(K)
    A = (rng.standard_normal((K, M) if ta else (M, K)) * s).astype(np.float32)
    B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    for mode in tc_modes():
        got = ops.gemm(to_dev(A), to_dev(B), transA=ta, transB=tb, mode=mode).cpu().numpy()
        ref = ref_gemm(A, B, ta, tb)
        # FFMA: fp32 round-to-nearest.  tcgen05 3xTF32: the tensor core accumulates with
        # round-toward-zero -> error ~ 3e-5 of the output magnitude per 600-long chain
        atol = 5e-6 if mode == "fma" else 4e-5 * max(1.0, np.abs(ref).max())
        assert_close(got, ref, atol=atol, what="%s %s" % (mode, (ta, tb, M, N, K)))


def test_gemm_epilogues_beta_bias_act_mask():
    from graphconvgeo_b200 import ops
    rng = np.random.RandomState(0)
    M, N, K = 300, 200, 150
    A = (rng.standard_normal((M, K)) / np.sqrt(K)).astype(np.float32)
    B = rng.standard_normal((K, N)).astype(np.float32)
    C0 = rng.standard_normal((M, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    mask = rng.standard_normal((M, N)).astype(np.float32)
    for mode in tc_modes():
        out = to_dev(C0)
        ops.gemm(to_dev(A), to_dev(B), out=out, beta=1.0, bias=to_dev(bias), act="sigmoid", mode=mode)
        ref = 1.0 / (1.0 + np.exp(-(A.astype(np.float64) @ B + C0 + bias)))
        tol = 1.0 if mode == "fma" else 20.0
        assert_close(out.cpu().numpy(), ref, atol=2e-6 * tol, what=mode)
        out = ops.gemm(to_dev(A), to_dev(B), mask=to_dev(mask), mask_act="rectify", mode=mode)
        assert_close(out.cpu().numpy(), (A.astype(np.float64) @ B) * (mask > 0), atol=5e-6 * tol, what=mode)
        out = ops.gemm(to_dev(A), to_dev(B), mask=to_dev(np.tanh(mask)), mask_act="tanh", mode=mode)
        assert_close(out.cpu().numpy(), (A.astype(np.float64) @ B) * (1 - np.tanh(mask) ** 2), atol=5e-6 * tol, what=mode)


@pytest.mark.parametrize("split", [0, 1, 3, 16])
def test_split_k_is_deterministic_and_correct(split):
    """dW = H^T.dZ: tall-skinny reduction over the node dimension."""
    from graphconvgeo_b200 import ops
    rng = np.random.RandomState(1)
    K, M, N = 20000, 96, 40
    A = (rng.standard_normal((K, M)) * 0.05).astype(np.float32)
    B = (rng.standard_normal((K, N)) * 0.05).astype(np.float32)
    Ad, Bd = to_dev(A), to_dev(B)
    g1 = ops.gemm(Ad, Bd, transA=True, split_k=split, mode="fma").cpu().numpy()
    g2 = ops.gemm(Ad, Bd, transA=True, split_k=split, mode="fma").cpu().numpy()
    assert np.array_equal(g1, g2)
    assert_close(g1, A.astype(np.float64).T @ B.astype(np.float64), atol=3e-6)


def test_unaligned_leading_dimensions():
    from graphconvgeo_b200 import ops
    rng = np.random.RandomState(2)
    A = torch.from_numpy((rng.standard_normal((50, 37)) * 0.1).astype(np.float32)).cuda()     # ld 37
    B = torch.from_numpy(rng.standard_normal((37, 29)).astype(np.float32)).cuda()             # ld 29
    out = torch.empty(50, 29, device="cuda")
    ops.gemm(A, B, out=out, mode="fma")
    assert_close(out.cpu().numpy(), A.cpu().numpy().astype(np.float64) @ B.cpu().numpy(), atol=3e-6)


def test_errors():
    from graphconvgeo_b200 import ops
    with pytest.raises(ValueError):
        ops.gemm(torch.zeros(4, 5, device="cuda"), torch.zeros(6, 3, device="cuda"))
    with pytest.raises(TypeError):
        ops.gemm(torch.zeros(4, 5), torch.zeros(5, 3))


def test_tensor_core_engine_is_really_used_and_long_contractions_are_chained():
    """the tcgen05 path must run (not silently fall back) for aligned operands; the weight-gradient
    shape (K = #nodes) is cut into <= 8192-long accumulation chains (deterministic split-K)."""
    from graphconvgeo_b200 import _lib, ops
    if "tf32x3" not in tc_modes():
        pytest.skip("tcgen05 engine unavailable")
    rng = np.random.RandomState(5)
    K, M, N = 150000, 300, 128
    A = (rng.standard_normal((K, M)) * 0.02).astype(np.float32)
    B = (rng.standard_normal((K, N)) * 0.02).astype(np.float32)
    Ad, Bd = to_dev(A), to_dev(B)
    n0 = ops.launch_count(reset=True)
    g1 = ops.gemm(Ad, Bd, transA=True, mode="tf32x3").cpu().numpy()
    launched = ops.launch_count(reset=True)
    assert launched == 4, launched        # 2 operand splits + tcgen05 kernel + split-K reduce
    g2 = ops.gemm(Ad, Bd, transA=True, mode="tf32x3").cpu().numpy()
    assert np.array_equal(g1, g2)
    ref = A.astype(np.float64).T @ B.astype(np.float64)
    assert_close(g1, ref, atol=2e-5 * np.abs(ref).max(), what="tf32x3 weight gradient")
    fma = ops.gemm(Ad, Bd, transA=True, mode="fma").cpu().numpy()
    assert_close(fma, ref, atol=2e-6 * np.abs(ref).max())


def test_presplit_operands_give_identical_results():
    """gcg_gemm_presplit_f32: an operand split once and shared by several GEMMs == splitting inside each call."""
    from graphconvgeo_b200 import ops
    if "tf32x3" not in tc_modes():
        pytest.skip("tcgen05 engine unavailable")
    rng = np.random.RandomState(9)
    A = to_dev((rng.standard_normal((3000, 600)) * 0.05).astype(np.float32))
    B = to_dev((rng.standard_normal((600, 256)) * 0.05).astype(np.float32))
    D = to_dev((rng.standard_normal((3000, 256)) * 0.05).astype(np.float32))
    sA, sD = ops.tf32_split(A), ops.tf32_split(D)
    assert sA.valid and sD.valid
    n0 = ops.launch_count(reset=True)
    r1 = ops.gemm(A, B, mode="tf32x3", a_split=sA)
    assert ops.launch_count(reset=True) == 2            # split of B + tcgen05 kernel; A's split was reused
    r0 = ops.gemm(A, B, mode="tf32x3")
    assert ops.launch_count(reset=True) == 3
    assert torch.equal(r0, r1)
    t1 = ops.gemm(A, D, transA=True, mode="tf32x3", a_split=sA, b_split=sD)
    t0 = ops.gemm(A, D, transA=True, mode="tf32x3")
    assert torch.equal(t0, t1)
    # a split of a different tensor is ignored, not misused
    other = ops.tf32_split(to_dev(np.zeros((3000, 600), np.float32)))
    assert torch.equal(ops.gemm(A, B, mode="tf32x3", a_split=other), r0)
    # the FFMA engine ignores pre-split operands
    assert_close(ops.gemm(A, B, mode="fma", a_split=sA).cpu().numpy(), r0.cpu().numpy(), atol=1e-5)


def test_opt_in_single_pass_tf32_for_long_weight_gradient_contractions(monkeypatch):
    """GCG_GEMM_LONGK_TF32 (off by default): TN products over K >= K0 rows whose operands both carry round-to-nearest
    hi copies run one tf32 pass on those copies.  Off: bit-identical to 3xTF32.  On: unbiased (the hardware would
    TRUNCATE raw fp32 inputs: a relative bias of ~5e-4 with same-sign products) and, because the input rounding
    averages over K, as close to float64 as the 3-pass product at this K."""
    from graphconvgeo_b200 import ops
    if "tf32x3" not in tc_modes():
        pytest.skip("tcgen05 engine unavailable")
    rng = np.random.RandomState(11)
    K, M, N = 200000, 256, 384
    A = np.abs(rng.standard_normal((K, M)) * 0.5).astype(np.float32)       # same-sign products: a bias would show
    B = np.abs(rng.standard_normal((K, N)) * 1e-3).astype(np.float32)
    Ad, Bd = to_dev(A), to_dev(B)
    sA, sB = ops.tf32_split(Ad), ops.tf32_split(Bd)
    ref = A.astype(np.float64).T @ B.astype(np.float64)
    assert ops._LONGK_TF32 == 0
    x3 = ops.gemm(Ad, Bd, transA=True, a_split=sA, b_split=sB, mode="tf32x3")
    off = ops.gemm(Ad, Bd, transA=True, a_split=sA, b_split=sB)
    assert torch.equal(x3, off)
    monkeypatch.setattr(ops, "_LONGK_TF32", 65536)
    on = ops.gemm(Ad, Bd, transA=True, a_split=sA, b_split=sB)
    assert not torch.equal(on, off)
    e_on = (on.cpu().numpy() - ref) / ref
    e_x3 = (x3.cpu().numpy() - ref) / ref
    # the single pass IS the hi.hi part of the 3-pass product (same chains); what it drops are the lo cross terms,
    # zero-mean after round-to-nearest and averaged over K: ~4e-4 / sqrt(K) = 9e-7 relative per output
    d = e_on - e_x3
    assert abs(d.mean()) < 1e-6, d.mean()                        # unbiased (truncated inputs: ~ -5e-4, below)
    assert np.abs(d).max() < 1e-5, np.abs(d).max()
    # the rule needs BOTH rounded copies and a TN product; otherwise nothing changes
    assert torch.equal(ops.gemm(Ad, Bd, transA=True, a_split=sA), ops.gemm(Ad, Bd, transA=True, a_split=sA, mode="tf32x3"))
    raw = ops.gemm(Ad, Bd, transA=True, mode="tf32").cpu().numpy()         # raw fp32 into the tensor core: truncation
    assert ((raw - ref) / ref - e_x3).mean() < -1e-4


@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("M,N,K,chained", [(1024, 600, 600, False), (1100, 256, 320, True), (640, 1024, 1000, False),
                                           (129, 70, 64, True), (2000, 600, 8200, False)])
def test_clustered_multicast_gemm_equals_single_cta(ta, tb, M, N, K, chained, monkeypatch):
    """2-CTA clusters (the B tile multicast into both CTAs, odd M-tile counts, all four operand majors, chained
    accumulation, split-K) must give the very bits of the single-CTA schedule: same MMAs in the same order."""
    from graphconvgeo_b200 import ops
    if "tf32x3" not in tc_modes():
        pytest.skip("tcgen05 engine unavailable")
    rng = np.random.RandomState(M + N + K)
    s = 1.0 / np.sqrt(K)
    A = to_dev((rng.standard_normal((K, M) if ta else (M, K)) * s).astype(np.float32))
    B = to_dev(rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32))
    bias = to_dev(rng.standard_normal(N).astype(np.float32))
    monkeypatch.setenv("GCG_GEMM_CLUSTER", "1")
    one = ops.gemm(A, B, transA=ta, transB=tb, mode="tf32x3", bias=bias, act="rectify", chained=chained).clone()
    monkeypatch.setenv("GCG_GEMM_CLUSTER", "2")
    monkeypatch.setenv("GCG_GEMM_CLUSTER_MIN_TILES", "1")
    two = ops.gemm(A, B, transA=ta, transB=tb, mode="tf32x3", bias=bias, act="rectify", chained=chained)
    torch.cuda.synchronize()
    assert torch.equal(one, two), (ta, tb, M, N, K, float((one - two).abs().max()))
    ref = np.maximum(ref_gemm(A.cpu().numpy(), B.cpu().numpy(), ta, tb) + bias.cpu().numpy()[None, :], 0)
    assert_close(two.cpu().numpy(), ref, atol=4e-5 * max(1.0, np.abs(ref).max()))
