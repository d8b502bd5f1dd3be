"""Plain-Python restatement of the reference's mention-graph projection (TEST INFRASTRUCTURE).

Follows data.py:226-250 (``efficient_collaboration_weighted_projected_graph2``) and the celebrity
filter of data.py:364-370 on adjacency *sets* instead of a networkx graph.

Parity status: PINNED -- tests/golden/projection_golden.npz holds inputs and outputs of the reference's
own function (its source is loaded from /root/reference/data.py by tests/golden/make_projection_golden.py
and run on networkx graphs); this restatement and the CUDA path must both reproduce them exactly.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def adjacency_sets(B):
    B = sp.csr_matrix(B)
    return [set(B.indices[B.indptr[i]:B.indptr[i + 1]].tolist()) for i in range(B.shape[0])]


def remove_celebrities(adj, n_targets, celebrity_threshold):
    """data.py:364-370: non-target nodes with degree 1 or above the threshold are removed."""
    celebrities = [i for i in range(n_targets, len(adj))
                   if len(adj[i]) == 1 or len(adj[i]) > celebrity_threshold]
    gone = set(celebrities)
    out = []
    for i, nb in enumerate(adj):
        out.append(set() if i in gone else {j for j in nb if j not in gone})
    return out


def project(adj, n_targets):
    """data.py:226-250.  Returns the sorted edge list [(n1, n2)] with n1 < n2 over target users."""
    edges = set()
    for m, nbrs in enumerate(adj):
        target_nbrs = [t for t in nbrs if t < n_targets]                 # :238
        if m < n_targets:                                                # :239-243
            for n in target_nbrs:
                if m < n:
                    edges.add((m, n))
        for n1 in target_nbrs:                                           # :244-248
            for n2 in target_nbrs:
                if n1 < n2:
                    edges.add((n1, n2))
    return sorted(edges)


def edges_to_csr(edges, n):
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    a = sp.csr_matrix((np.ones(len(e), np.float32), (e[:, 0], e[:, 1])), shape=(n, n))
    a = sp.csr_matrix(((a + a.T) > 0).astype(np.float32))
    a.sort_indices()
    return a
