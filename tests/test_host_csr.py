"""Host-side C++ CSR helpers of libgcg.so (no GPU): transpose, row gather, node permutation, column-block
split -- property tests (hypothesis) against scipy on random CSR matrices with empty rows, duplicate row
selections and hub rows, as SURVEY.md section 4 asks."""
import numpy as np
import pytest
import scipy.sparse as sp

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

from util import random_csr  # noqa: E402


def _csr(m, built_lib):
    from graphconvgeo_b200.sparse import CSRMatrix
    return CSRMatrix.from_scipy(m, device="cpu")


def _eq(got, ref):
    ref = sp.csr_matrix(ref)
    ref.sort_indices()
    g = got.to_scipy()
    assert g.shape == ref.shape
    assert np.array_equal(g.indptr, ref.indptr) and np.array_equal(g.indices, ref.indices)
    assert np.array_equal(g.data, ref.data.astype(np.float32))


shapes = st.tuples(st.integers(1, 60), st.integers(1, 70), st.integers(0, 9), st.integers(0, 2 ** 31 - 1))


@settings(max_examples=40, deadline=None)
@given(shapes)
def test_transpose_is_scipy_transpose(built_lib, shape):
    n, m, deg, seed = shape
    rng = np.random.RandomState(seed)
    a = random_csr(rng, n, m, deg, hub_rows=(0,), hub_deg=m, empty_frac=0.3)
    _eq(_csr(a, built_lib).T, a.T)
    _eq(_csr(a, built_lib).T.T, a)


@settings(max_examples=40, deadline=None)
@given(shapes, st.integers(0, 80))
def test_gather_rows_with_duplicates(built_lib, shape, n_sel):
    n, m, deg, seed = shape
    rng = np.random.RandomState(seed)
    a = random_csr(rng, n, m, deg, empty_frac=0.3)
    idx = rng.randint(0, n, size=n_sel).astype(np.int32)          # duplicates on purpose (tensormain.py:226)
    _eq(_csr(a, built_lib).gather_rows(idx), a[idx] if n_sel else sp.csr_matrix((0, m), dtype=np.float32))


@settings(max_examples=40, deadline=None)
@given(st.tuples(st.integers(1, 50), st.integers(0, 8), st.integers(0, 2 ** 31 - 1)))
def test_symmetric_permutation_keeps_the_graph(built_lib, shape):
    n, deg, seed = shape
    rng = np.random.RandomState(seed)
    a = random_csr(rng, n, n, deg, empty_frac=0.2)
    a = sp.csr_matrix(a + a.T)
    order = rng.permutation(n).astype(np.int32)
    inv = np.empty(n, np.int32)
    inv[order] = np.arange(n, dtype=np.int32)
    got = _csr(a, built_lib).permute(order, col_map=inv)
    _eq(got, a[order][:, order])
    # rows only (X is permuted like this: mlpconv reorder)
    _eq(_csr(a, built_lib).permute(order), a[order])


@settings(max_examples=25, deadline=None)
@given(st.tuples(st.integers(2, 40), st.integers(8, 90), st.integers(1, 12), st.integers(0, 2 ** 31 - 1)),
       st.integers(1, 30))
def test_column_block_split_partitions_the_selected_rows(built_lib, shape, block_cols):
    import ctypes as C
    from graphconvgeo_b200 import _lib
    n, m, deg, seed = shape
    rng = np.random.RandomState(seed)
    a = random_csr(rng, n, m, deg, hub_rows=(1,), hub_deg=m, empty_frac=0.2)
    sel = np.unique(rng.randint(0, n, size=max(1, n // 2))).astype(np.int32)
    nb = -(-m // block_cols)
    ip = a.indptr.astype(np.int32)
    ix = a.indices.astype(np.int32)
    d = a.data.astype(np.float32)
    ptr = lambda x: x.ctypes.data_as(C.c_void_p)
    o_ip = np.empty(nb * (len(sel) + 1), np.int32)
    o_off = np.empty(nb + 1, np.int64)
    L = _lib.lib()
    _lib.check(L.gcg_csr_split_colblocks_host(ptr(ip), ptr(ix), ptr(d), ptr(sel), len(sel), block_cols, nb,
                                              ptr(o_ip), ptr(o_off), None, None), "size")
    o_ix = np.empty(int(o_off[-1]), np.int32)
    o_d = np.empty(int(o_off[-1]), np.float32)
    _lib.check(L.gcg_csr_split_colblocks_host(ptr(ip), ptr(ix), ptr(d), ptr(sel), len(sel), block_cols, nb,
                                              ptr(o_ip), ptr(o_off), ptr(o_ix), ptr(o_d)), "fill")
    total = sp.csr_matrix((len(sel), m), dtype=np.float32)
    for b in range(nb):
        s, e = int(o_off[b]), int(o_off[b + 1])
        blk = sp.csr_matrix((o_d[s:e], o_ix[s:e], o_ip[b * (len(sel) + 1):(b + 1) * (len(sel) + 1)]),
                            shape=(len(sel), m))
        if blk.nnz:
            assert blk.indices.min() >= b * block_cols and blk.indices.max() < (b + 1) * block_cols
        total = total + blk
    ref = a[sel]
    assert (abs(total - ref)).nnz == 0 and total.nnz == ref.nnz


def test_graph_helpers_on_the_host():
    from graphconvgeo_b200.graph import mention_incidence, remove_celebrities
    # 3 users (0..2), names 3..5: name 3 mentioned once (dropped), name 4 by all three, name 5 by two
    e = [(0, 3), (0, 4), (1, 4), (2, 4), (1, 5), (2, 5), (0, 0), (1, 1), (2, 2)]
    B = sp.csr_matrix((np.ones(len(e)), ([a for a, _ in e], [b for _, b in e])), shape=(6, 6))
    B = sp.csr_matrix(((B + B.T) > 0).astype(np.float32))
    f = remove_celebrities(B, 3, 2)                                # data.py:364-370: deg 1 or > 2 go
    assert f[3].nnz == 0 and f[4].nnz == 0 and f[5].nnz == 2 and f[:, 3].nnz == 0
    R = mention_incidence(f, 3)
    G = (R.T @ R).toarray() > 0
    np.fill_diagonal(G, False)
    assert G.tolist() == [[False, False, False], [False, False, True], [False, True, False]]
