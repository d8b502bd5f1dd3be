"""The @-mention graph -> user-user graph step that produces the GCN's adjacency.

Reference: ``DataLoader.get_graph`` (data.py:302-375) builds a graph over target users (ids 0..N-1, each with a
self loop, data.py:309-310) and the other mentioned names (ids >= N), removes "celebrity" names
(data.py:364-370) and projects it onto the target users with
``efficient_collaboration_weighted_projected_graph2`` (data.py:226-250): for EVERY node m, all target
neighbours of m (and m itself when it is a target) become pairwise connected, unweighted.

With R[m, t] = 1 iff target t is a neighbour of node m or t == m, that edge set is exactly the off-diagonal
pattern of R^T R -- a boolean sparse x sparse product, done on the GPU by the pattern mode of the SpGEMM
kernel (gcg_spgemm_count_csr / gcg_spgemm_fill_pattern_csr).  The text side (regex over tweets,
data.py:311-362) is data loading and stays out of scope; the input here is the mention graph's adjacency.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .sparse import CSRMatrix, spgemm_pattern


def remove_celebrities(B, n_targets, celebrity_threshold):
    """data.py:364-370: drop every non-target node whose degree is 1 or exceeds the threshold.
    ``B``: symmetric scipy adjacency [M, M]; returns the adjacency with those rows/columns emptied
    (node ids are kept, as removing a node from the networkx graph does not renumber the others)."""
    B = sp.csr_matrix(B)
    deg = np.diff(B.indptr)
    drop = np.zeros(B.shape[0], bool)
    drop[n_targets:] = (deg[n_targets:] == 1) | (deg[n_targets:] > celebrity_threshold)
    keep = sp.diags((~drop).astype(B.dtype))
    out = sp.csr_matrix(keep @ B @ keep)
    out.eliminate_zeros()
    return out


def mention_incidence(B, n_targets):
    """R [M, N]: R[m, t] = 1 iff target t is adjacent to node m, plus R[t, t] = 1 (the self loops of
    data.py:309-310; adding them when absent does not change the projected edge set)."""
    B = sp.csr_matrix(B)
    M = B.shape[0]
    R = sp.csr_matrix(B[:, :n_targets], dtype=np.float32)
    eye = sp.csr_matrix((np.ones(n_targets, np.float32), (np.arange(n_targets), np.arange(n_targets))),
                        shape=(M, n_targets))
    R = sp.csr_matrix(R + eye)
    R.data[:] = 1.0
    R.sort_indices()
    return R


def project_mention_graph(B, n_targets, celebrity_threshold=None, device="cuda") -> CSRMatrix:
    """The user-user graph of data.py:364-373 as a device CSR pattern [N, N] (binary, symmetric, no self
    edges, ascending columns) -- ready for ``sparse.build_ahat_device(indptr, indices, N)``."""
    if celebrity_threshold is not None:
        B = remove_celebrities(B, n_targets, celebrity_threshold)
    R = mention_incidence(B, n_targets)
    Rd = CSRMatrix.from_scipy(R, device=device)
    RT = CSRMatrix.from_scipy(sp.csr_matrix(R.T), device=device)
    return spgemm_pattern(RT, Rd, drop_diagonal=True)
