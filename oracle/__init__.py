"""CPU oracle for the graphconvgeo GCN propagation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``graphconvgeo_b200`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
or the CPU comparator, never as the product path.

Parity status (see DESIGN.md "Oracle"):

* GCN layers / loss / Adam / fit loop (``gcn_oracle.py``): **pinned** to outputs of the
  reference's own source executed under stub theano / lasagne modules
  (tests/golden/make_layers_golden.py: forward bodies, A_hat, geo_eval, bit for bit;
  tests/golden/make_fit_golden.py: a whole ``MLPCONV.fit`` run).  The hand-written backward of
  the >2-layer / gated extension is cross-checked against torch-CPU autograd.
* Highway gate: **parity unpinned** -- the gate is not in the reference at all;
  the oracle restates the formula given in BASELINE.json's ``north_star``.
* Input smoothing ``X_conv = H * X`` (``mlp_oracle.smooth_features``): **pinned** -- the
  function calls scipy's ``csr_matmat`` + ``astype`` exactly as main.py:528-530 does, i.e. it
  executes the reference's own third-party routine; the CUDA SpGEMM is compared bit for bit.
  The minibatch MLP around it (``mlp_oracle.MLPOracle`` / ``fit``, mlp.py:121-314) is
  **unpinned** like the GCN and cross-checked against torch-CPU autograd.
* Mention-graph projection (``graph_oracle.py``, data.py:226-250,364-370): **pinned** -- equal to
  golden graphs produced by the reference's own function (tests/golden/make_projection_golden.py).
* kd-tree labels (``kdtree_oracle.py``): **pinned** -- checked bit-for-bit
  against the reference's own ``kdtree.py`` (imported from /root/reference by
  ``tests/golden/make_kdtree_golden.py``, outputs committed under
  ``tests/golden/``).
"""
