"""the @-mention graph -> user-user graph step (data.py:226-250, 364-373).
CPU part: the plain-Python oracle and the R^T R formulation the GPU path uses, against golden vectors made
by the reference's own function (tests/golden/make_projection_golden.py).  GPU part: the CUDA pattern SpGEMM
against the same vectors, and the projected graph feeding the device A_hat builder."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import gcn_oracle as go
from oracle import graph_oracle as gro

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "projection_golden.npz")
CASES = ["small", "names_only", "direct_only", "celebrity_heavy", "sparse"]


def load(case):
    z = np.load(GOLDEN)
    n, M, thr = (int(v) for v in z[case + "__meta"])
    e = z[case + "__edges"]
    B = sp.csr_matrix((np.ones(len(e), np.float32), (e[:, 0], e[:, 1])), shape=(M, M))
    B = sp.csr_matrix(((B + B.T) > 0).astype(np.float32))
    return B, n, thr, z[case + "__proj_raw"], z[case + "__proj_filtered"]


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_the_reference_projection(case):
    B, n, thr, raw, filt = load(case)
    adj = gro.adjacency_sets(B)
    assert np.array_equal(np.array(gro.project(adj, n)).reshape(-1, 2), raw)
    adj_f = gro.remove_celebrities(adj, n, thr)
    assert np.array_equal(np.array(gro.project(adj_f, n)).reshape(-1, 2), filt)


@pytest.mark.parametrize("case", CASES)
def test_incidence_product_formulation_on_the_host(case):
    """offdiag(pattern(R^T R)) == the reference's edge set: the identity the GPU path relies on."""
    from graphconvgeo_b200.graph import mention_incidence, remove_celebrities
    B, n, thr, raw, filt = load(case)
    for Bx, ref in ((B, raw), (remove_celebrities(B, n, thr), filt)):
        R = mention_incidence(Bx, n)
        G = sp.csr_matrix(R.T @ R)
        G.setdiag(0)
        G.eliminate_zeros()
        G.sort_indices()
        want = gro.edges_to_csr(ref, n)
        assert np.array_equal(G.indptr, want.indptr) and np.array_equal(G.indices, want.indices)


def test_projection_without_self_loops_is_the_same_graph():
    """get_graph adds target self loops (data.py:309-310); the projection does not depend on them."""
    B, n, thr, raw, _ = load("small")
    Bn = B.tolil()
    Bn.setdiag(0)
    Bn = sp.csr_matrix(Bn)
    Bn.eliminate_zeros()
    assert np.array_equal(np.array(gro.project(gro.adjacency_sets(Bn), n)).reshape(-1, 2), raw)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_projection_matches_the_reference(case):
    import torch
    from graphconvgeo_b200.graph import project_mention_graph
    from graphconvgeo_b200.sparse import build_ahat_device
    B, n, thr, raw, filt = load(case)
    for kw, ref in ((dict(), raw), (dict(celebrity_threshold=thr), filt)):
        G = project_mention_graph(B, n, **kw)
        torch.cuda.synchronize()
        want = gro.edges_to_csr(ref, n)
        assert G.shape == (n, n)
        assert np.array_equal(G.indptr.cpu().numpy(), want.indptr)
        assert np.array_equal(G.indices.cpu().numpy(), want.indices)
    # the projected graph feeds the device A_hat builder: same bits as the oracle's normalisation
    a = build_ahat_device(G.indptr, G.indices, n)
    torch.cuda.synchronize()
    ref_a = go.build_ahat(want)
    assert np.array_equal(a.indices.cpu().numpy(), ref_a.indices) and np.array_equal(a.data.cpu().numpy(), ref_a.data)


@pytest.mark.gpu
def test_gpu_pattern_product_general_shapes():
    import torch
    from graphconvgeo_b200.sparse import CSRMatrix, spgemm_pattern
    from util import random_csr
    rng = np.random.RandomState(6)
    A = random_csr(rng, 300, 1100, 9, hub_rows=(7,), hub_deg=700)
    Bm = random_csr(rng, 1100, 40000, 12)
    ref = sp.csr_matrix(abs(A) @ abs(Bm))
    ref.sort_indices()
    got = spgemm_pattern(CSRMatrix.from_scipy(A), CSRMatrix.from_scipy(Bm))
    torch.cuda.synchronize()
    assert np.array_equal(got.indptr.cpu().numpy(), ref.indptr) and np.array_equal(got.indices.cpu().numpy(), ref.indices)
    sq = random_csr(rng, 500, 500, 6)
    ref = sp.csr_matrix(abs(sq) @ abs(sq))
    ref.setdiag(0)
    ref.eliminate_zeros()
    ref.sort_indices()
    got = spgemm_pattern(CSRMatrix.from_scipy(sq), CSRMatrix.from_scipy(sq), drop_diagonal=True)
    torch.cuda.synchronize()
    assert np.array_equal(got.indptr.cpu().numpy(), ref.indptr) and np.array_equal(got.indices.cpu().numpy(), ref.indices)
