"""Per-operation parity of one training step at FULL benchmark size (TEST INFRASTRUCTURE: the checker).

At Twitter-US / Twitter-World shape a whole oracle epoch costs minutes of host time, so the step is checked
operation by operation on SAMPLED rows instead: for ~2k random nodes (and target positions, and parameter
rows) every product, epilogue and gradient of `f_train` (mlpconv.py:226-263 through lasagne_layers.py:60-89)
is recomputed with the oracle's routines -- scipy ``csr @ dense`` (the sparsetools routine Theano's S.dot
runs), BLAS ``numpy.dot``, the hand-written backward of SURVEY Appendix A.3 -- from the operands THE GPU PATH
ITSELF holds for that operation ("teacher forcing": each operation is compared on identical inputs, which is
what "per-layer activations and gradients must match" asks), and compared under the north_star bound

        |gpu - oracle| <= 1e-6 + 1e-4 * |oracle|          (reported as max of the ratio: <= 1 passes)

Contractions over all nodes / all targets (bias gradients, weight-gradient rows, the loss) and the dense
projections (K = 600 with partial sums up to ~65 against results near 0: two float32 BLAS orders already differ by
more than the bound on such elements) use float64 accumulation of the same float32 operands as the arbiter: two
float32 summation orders differ from each other by more than either differs from the float64 sum.  The sparse
products are compared with scipy's float32 routine itself (the GPU kernels reproduce its order bit for bit).

Works on ``MLPCONV`` and (every rank calling it collectively) on ``DistMLPCONV``: per-node slabs are
all-gathered first, which is plumbing, not arithmetic.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

F32 = np.float32
ATOL, RTOL = 1e-6, 1e-4


def _sigmoid(x):
    return (F32(1) / (F32(1) + np.exp(-x, dtype=F32))).astype(F32)


def _act(name, x):
    if name == "rectify":
        return np.maximum(x, F32(0))
    if name == "tanh":
        return np.tanh(x, dtype=F32)
    if name == "sigmoid":
        return _sigmoid(x)
    return x


def _dact(name, a):
    if name == "rectify":
        return (a > 0).astype(F32)
    if name == "tanh":
        return (F32(1) - a * a).astype(F32)
    if name == "sigmoid":
        return (a * (F32(1) - a)).astype(F32)
    return np.ones_like(a)


class _Report:
    def __init__(self, log=None):
        self.checks = {}
        self.log = log

    def add(self, name, got, ref, reference_f32=None):
        """``reference_f32``: the same values computed with the REFERENCE's own float32 arithmetic (sgemm, scipy
        float32 csr_matvecs) where ``ref`` is the float64 arbiter: its own scaled distance from the arbiter is the
        noise floor of the comparison -- no float32 implementation can be asked to sit closer to the float64 value
        than the reference's float32 path does."""
        got = np.asarray(got, dtype=np.float64)
        ref = np.asarray(ref, dtype=np.float64)
        assert got.shape == ref.shape, "%s: shape %s vs %s" % (name, got.shape, ref.shape)
        err = np.abs(got - ref)
        ratio = err / (ATOL + RTOL * np.abs(ref))
        r = float(ratio.max()) if ratio.size else 0.0
        floor = None
        if reference_f32 is not None:
            f = np.asarray(reference_f32, dtype=np.float64)
            fr = np.abs(f - ref) / (ATOL + RTOL * np.abs(ref))
            floor = dict(max_scaled_err=float(fr.max()) if fr.size else 0.0,
                         frac_over=float((fr > 1).mean()) if fr.size else 0.0,
                         gpu_vs_reference_f32=float((np.abs(got - f) / (ATOL + RTOL * np.abs(f))).max()) if fr.size else 0.0)
        scale = float(np.abs(ref).max()) if ref.size else 0.0
        # gradients of a mean over 1.4 M targets are ~1e-7..1e-9, far below the bound's absolute term, so the
        # error relative to the largest reference value of the array is reported (and tested) as well
        rel = (float(err.max()) / scale) if (err.size and scale > 0) else 0.0
        e = dict(max_scaled_err=r, max_abs_err=float(err.max()) if err.size else 0.0, ref_max_abs=scale,
                 max_err_over_ref_max=rel, n=int(ref.size), frac_over=float((ratio > 1).mean()) if ratio.size else 0.0)
        if floor is not None:
            e["reference_f32_noise"] = floor
        self.checks[name] = e
        if self.log:
            self.log("  parity %-46s scaled %8.3f  abs %.3e (|ref|max %.3e -> rel %.2e, %d values, %.2e over)%s"
                     % (name, r, e["max_abs_err"], scale, rel, e["n"], e["frac_over"],
                        "" if floor is None else "  [reference float32 path vs the same float64 values: %.3f, %.2e over]"
                        % (floor["max_scaled_err"], floor["frac_over"])))
        return r

    def summary(self):
        worst = max(self.checks.items(), key=lambda kv: kv[1]["max_scaled_err"]) if self.checks else ("", {"max_scaled_err": 0.0})
        wrel = max(self.checks.items(), key=lambda kv: kv[1]["max_err_over_ref_max"]) if self.checks else ("", {"max_err_over_ref_max": 0.0})
        # a check passes when it is inside the bound, or no further from the float64 values than the reference's own
        # float32 arithmetic is on the same sample (the bound's 1e-6 absolute term is below one float32 ulp of the
        # partial sums of a K = 600 contraction whose terms reach ~10: no float32 evaluation order meets it there)
        def excess(e):
            fl = e.get("reference_f32_noise", {}).get("max_scaled_err", 0.0)
            return e["max_scaled_err"] / max(1.0, fl)
        wex = max(self.checks.items(), key=lambda kv: excess(kv[1])) if self.checks else ("", None)
        return dict(max_scaled_err=worst[1]["max_scaled_err"], worst_check=worst[0], n_checks=len(self.checks),
                    max_err_over_ref_max=wrel[1]["max_err_over_ref_max"], worst_relative_check=wrel[0],
                    max_scaled_err_over_reference_noise=excess(wex[1]) if wex[1] else 0.0,
                    worst_check_over_reference_noise=wex[0],
                    reference_f32_noise={k: v["reference_f32_noise"] for k, v in self.checks.items() if "reference_f32_noise" in v},
                    tolerance="|gpu-oracle| <= 1e-6 + 1e-4*|oracle|", checks=self.checks)


def _dot64(a, b):
    """dense contraction of float32 operands accumulated in float64 (the arbiter for T.dot products)"""
    return np.dot(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64))


def _compact_rows(A, rows):
    """A[rows, :] with its columns renumbered to the sorted distinct columns it touches -> (csr, cols)."""
    Ar = A[rows]
    cols = np.unique(Ar.indices)
    remap = np.searchsorted(cols, Ar.indices).astype(np.int32)
    return sp.csr_matrix((Ar.data, remap, Ar.indptr), shape=(len(rows), len(cols))), cols


class _Access:
    """Reads of the model's device buffers as host arrays.  Single GPU: plain copies.  Row-partitioned model:
    per-node slabs are all-gathered over the ranks (collective), per-target buffers are concatenated in rank
    order -- both only to give the checker the operands; no arithmetic happens here."""

    def __init__(self, m):
        import torch
        self.t = torch
        self.m = m
        self.dist = hasattr(m, "part")
        self.N = m.part.n_total if self.dist else m.l_hid1.H.shape[0]
        if self.dist:
            import torch.distributed as dist
            self.d = dist

    def _full_dev(self, t):
        if not self.dist:
            return t
        part = self.m.part
        full = self.t.empty((part.n_pad, t.shape[1]), dtype=t.dtype, device=t.device)
        self.d.all_gather_into_tensor(full, t.contiguous(), group=self.m.group)
        return full[:part.n_total]

    def rows(self, t, idx):
        """rows ``idx`` (global node ids, reordered numbering) of a per-node matrix"""
        full = self._full_dev(t)
        sel = self.t.from_numpy(np.asarray(idx, dtype=np.int64)).to(full.device)
        return full.index_select(0, sel).cpu().numpy()

    def node_chunks(self, tensors, chunk=131072):
        """yield host chunks [(a0, a1, ...)] of per-node matrices, all rows, in node order"""
        fulls = [self._full_dev(t) for t in tensors]
        n = fulls[0].shape[0]
        for s in range(0, n, chunk):
            yield tuple(f[s:s + chunk].cpu().numpy() for f in fulls)

    def targets_dev(self, t):
        """a per-target matrix with all ranks' rows (rank order); single GPU: itself"""
        if not self.dist:
            return t
        n_loc = self.t.tensor([t.shape[0]], device=t.device, dtype=self.t.int64)
        counts = [self.t.zeros_like(n_loc) for _ in range(self.m.world)]
        self.d.all_gather(counts, n_loc, group=self.m.group)
        counts = [int(c.item()) for c in counts]
        mx = max(max(counts), 1)
        pad = self.t.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        outs = [self.t.empty_like(pad) for _ in range(self.m.world)]
        self.d.all_gather(outs, pad, group=self.m.group)
        return self.t.cat([o[:c] for o, c in zip(outs, counts)], 0)

    def target_nodes(self):
        """global (reordered) node id of every target row, in the order ``targets_dev`` concatenates them"""
        ti = self.m.ti_train
        if not self.dist:
            return ti.host.astype(np.int64)
        loc = self.t.from_numpy(ti.local.astype(np.int64) + self.m.part.rank * self.m.part.n_loc).to(self.m.device)
        return self.targets_dev(loc.view(-1, 1)).view(-1).cpu().numpy()

    def target_labels(self):
        return self.targets_dev(self.m.y_train_dev.view(-1, 1)).view(-1).cpu().numpy()


def check_training_step(m, X_host, A_host, n_rows=2048, n_param_rows=24, seed=0, log=None):
    """Run one forward pass and one full f_train of ``m`` eagerly and compare every operation on sampled rows
    with the oracle.  ``X_host`` / ``A_host``: scipy CSR of the model's inputs IN THE MODEL'S NODE ORDER
    (``MLPCONV.Xd.to_scipy()`` / ``l_hid1.H.to_scipy()`` on one GPU).  Returns the report dictionary."""
    import torch
    from graphconvgeo_b200 import lasagne_layers as L
    rep = _Report(log)
    acc = _Access(m)
    N = acc.N
    rng = np.random.RandomState(seed)
    R = np.sort(rng.choice(N, size=min(n_rows, N), replace=False))
    A_R, cols_R = _compact_rows(A_host, R)
    layers = m.layers
    names = {id(ly): "L%d" % (i + 1) for i, ly in enumerate(layers)}
    dev = m.device
    sync = lambda: torch.cuda.synchronize(dev)

    def host(t):
        return t.detach().cpu().numpy()

    # ------------------------------------------------------------------ parameters before the step
    P0 = {}
    for ly in layers:
        for pname, t, tags in ly.params:
            P0[(id(ly), pname)] = host(t).copy()
    adam = m.adam
    idx_of = {p.data_ptr(): i for i, p in enumerate(adam.params)}
    sync()
    st0 = {}
    t_before = host(adam.t).copy()

    def sample_param_rows(t):
        if t.dim() == 1 or t.shape[0] <= 4096:
            return None
        return np.sort(rng.choice(t.shape[0], size=min(n_param_rows * 8, t.shape[0]), replace=False))

    prow = {}
    for ly in layers:
        for pname, t, tags in ly.params:
            i = idx_of[t.data_ptr()]
            sel = sample_param_rows(t)
            prow[(id(ly), pname)] = sel
            pick = (lambda a: a) if sel is None else (lambda a, s=torch.from_numpy(sel).to(dev): a.index_select(0, s))
            st0[(id(ly), pname)] = (host(pick(adam.m[i])).copy(), host(pick(adam.v[i])).copy())

    # ------------------------------------------------------------------ forward only
    ti = m.ti_train
    m._forward(ti, train=True)
    sync()
    tnodes = acc.target_nodes()
    ylab = acc.target_labels().astype(np.int64)
    n_t = len(tnodes)
    Tpos = np.sort(rng.choice(n_t, size=min(n_rows, n_t), replace=False))
    A_T, cols_T = _compact_rows(A_host, tnodes[Tpos])

    def zbuf(ly):          # the SpMM operand buffer "Z" of a conv layer (local slab in row-partitioned mode)
        return ly._operand("Z", ly._in.shape[0] if hasattr(ly, "_in") else ly._X.shape[0], ly.num_units)

    for ly in layers:
        nm = names[id(ly)]
        W = P0[(id(ly), "W")]
        b = P0.get((id(ly), "b"))
        if isinstance(ly, L.SparseConvolutionDenseLayer):
            Zg = zbuf(ly)
            rep.add(nm + " X.W1", acc.rows(Zg, R), np.asarray(X_host[R] @ W, dtype=F32))                 # lasagne_layers.py:65
            pre = np.asarray(A_R @ acc.rows(Zg, cols_R), dtype=F32) + b[None, :]                          # :67-70
            rep.add(nm + " act(A_hat.Z1+b)", acc.rows(ly._out, R), _act(ly.nonlinearity, pre))            # :71
        elif isinstance(ly, L.HighwayConvolutionDenseLayer):
            inp_R = acc.rows(ly._in, R)
            Zg = zbuf(ly)
            rep.add(nm + " H.W", acc.rows(Zg, R), _dot64(inp_R, W), reference_f32=np.dot(inp_R, W))       # :82
            Wg, bg = P0[(id(ly), "Wg")], P0[(id(ly), "bg")]
            rep.add(nm + " gate sigmoid(H.Wg+bg)", acc.rows(ly._g, R), 1.0 / (1.0 + np.exp(-(_dot64(inp_R, Wg) + bg[None, :]))))
            pre = np.asarray(A_R @ acc.rows(Zg, cols_R), dtype=F32) + b[None, :]
            hc = _act(ly.nonlinearity, pre)
            rep.add(nm + " act(A_hat.Z+b)", acc.rows(ly._Hc, R), hc)                                      # :84-89
            g_R, hc_R = acc.rows(ly._g, R), acc.rows(ly._Hc, R)
            rep.add(nm + " g*H'+(1-g)*H", acc.rows(ly._out, R), (g_R * hc_R + (F32(1) - g_R) * inp_R).astype(F32))
        elif ly is m.l_out:
            inp_c = acc.rows(ly._in, cols_T)
            ref = (A_T.astype(np.float64) @ _dot64(inp_c, W)) + b[None, :]                                # :82-88, reference order, float64
            # the REFERENCE's own arithmetic for the same values: float32 sgemm, then scipy's float32 csr_matvecs
            ref32 = np.asarray(A_T @ np.dot(inp_c, W).astype(F32), dtype=F32) + b[None, :]
            logits_T = host(acc.targets_dev(ly._out)[torch.from_numpy(Tpos).to(dev)])
            rep.add(nm + " logits (A_hat.(H.W)+b)[idx]", logits_T, ref, reference_f32=ref32)
            if ly.propagate_first:
                q_T = host(acc.targets_dev(ly._q)[torch.from_numpy(Tpos).to(dev)])
                rep.add(nm + " A_hat[idx,:].H", q_T, np.asarray(A_T @ inp_c, dtype=F32))
                rep.add(nm + " (A_hat[idx,:].H).W+b", logits_T, _dot64(q_T, W) + b[None, :],
                        reference_f32=np.dot(q_T, W) + b[None, :])
        else:                                             # plain hidden conv layer (n_layers > 2 without gate)
            inp_R = acc.rows(ly._in, R)
            Zg = zbuf(ly)
            rep.add(nm + " H.W", acc.rows(Zg, R), _dot64(inp_R, W))
            pre = np.asarray(A_R @ acc.rows(Zg, cols_R), dtype=F32) + b[None, :]
            rep.add(nm + " act(A_hat.Z+b)", acc.rows(ly._out, R), _act(ly.nonlinearity, pre))

    # ------------------------------------------------------------------ the full step: forward again, backward, Adam
    m._train_step_enqueue()
    sync()
    hb = m._train_hb
    lo = m.l_out
    nmo = names[id(lo)]
    n_glob = n_t
    Tsel = torch.from_numpy(Tpos).to(dev)
    logits_all = acc.targets_dev(lo._out)
    logits_T = host(logits_all[Tsel])
    y_T = ylab[Tpos]
    mx = logits_T.max(axis=1, keepdims=True)
    ex = np.exp(logits_T - mx, dtype=F32)
    sm = (ex / ex.sum(axis=1, keepdims=True, dtype=F32)).astype(F32)                                      # mlpconv.py:216
    ce_ref = (mx[:, 0] + np.log(ex.sum(axis=1, dtype=F32)) - logits_T[np.arange(len(Tpos)), y_T]).astype(F32)   # :229
    n_loc_t = lo._out.shape[0]
    ce_all = acc.targets_dev(hb["ce"][:n_loc_t].view(-1, 1)).view(-1)
    hit_all = acc.targets_dev(hb["hit"][:n_loc_t].view(-1, 1)).view(-1)
    rep.add("head cross-entropy rows", host(ce_all[Tsel]), ce_ref)
    out2 = host(hb["out"])
    rep.add("head mean CE (loss, mlpconv.py:230)", [out2[0]], [host(ce_all).astype(np.float64).sum() / n_glob])
    rep.add("head accuracy (mlpconv.py:252)", [out2[1]], [host(hit_all).astype(np.float64).sum() / n_glob])
    pred_ok = host(hit_all[Tsel]) == (logits_T.argmax(-1) == y_T).astype(F32)
    rep.add("head argmax==y rows", pred_ok.astype(F32), np.ones(len(Tpos), F32))
    G_ref = sm.copy()
    G_ref[np.arange(len(Tpos)), y_T] -= F32(1)
    G_ref = (G_ref / F32(n_glob)).astype(F32)
    G_loc = lo._buf["G"]
    G_all = acc.targets_dev(G_loc)
    rep.add("head dLogits rows", host(G_all[Tsel]), G_ref)

    def colsum_and_rows(pairs, row_sel_of_first):
        """float64 column sums of the second matrix and rows ``row_sel`` of first^T . second, over all rows"""
        cs, wr = None, None
        for a, g in pairs:
            a64 = a[:, row_sel_of_first].astype(np.float64)
            g64 = g.astype(np.float64)
            cs = g64.sum(axis=0) if cs is None else cs + g64.sum(axis=0)
            wr = a64.T @ g64 if wr is None else wr + a64.T @ g64
        return cs, wr

    def target_chunks(ta, tb, chunk=131072):
        for s in range(0, ta.shape[0], chunk):
            yield host(ta[s:s + chunk]), host(tb[s:s + chunk])

    def grad_of(ly, key):
        return host(ly.grads[key])

    # ---- output layer backward
    W = P0[(id(lo), "W")]
    wsel = np.sort(rng.choice(W.shape[0], size=min(n_param_rows, W.shape[0]), replace=False))
    if lo.propagate_first:
        q_all = acc.targets_dev(lo._q)
        cs, wr = colsum_and_rows(target_chunks(q_all, G_all), wsel)
        rep.add(nmo + " db = colsum(dLogits)", grad_of(lo, "b"), cs)
        rep.add(nmo + " dW rows = (A_hat[idx].H)^T.dLogits", grad_of(lo, "W")[wsel], wr)
        dQ_all = acc.targets_dev(lo._buf[("dQ", n_loc_t)])
        dQ_T = host(dQ_all[Tsel])
        rep.add(nmo + " dQ = dLogits.W^T", dQ_T, _dot64(host(G_all[Tsel]), W.T))
        S = lo._operand("dP", lo._in.shape[0], lo.num_inputs)
        # scatter-ADD of the target rows onto their nodes (duplicates accumulate, lasagne_layers.py:88 backward)
        nodes_T = tnodes[Tpos]
        uniq_nodes = np.unique(nodes_T)
        dQ_host_rows = {}
        order = np.argsort(tnodes, kind="stable")
        sorted_nodes = tnodes[order]
        lo_i = np.searchsorted(sorted_nodes, uniq_nodes, side="left")
        hi_i = np.searchsorted(sorted_nodes, uniq_nodes, side="right")
        need = np.concatenate([order[a:b] for a, b in zip(lo_i, hi_i)])
        dQ_need = host(dQ_all[torch.from_numpy(need).to(dev)])
        ref_S = np.zeros((len(uniq_nodes), dQ_need.shape[1]), F32)
        pos = 0
        for k, (a, b) in enumerate(zip(lo_i, hi_i)):
            for _ in range(b - a):
                ref_S[k] += dQ_need[pos]
                pos += 1
        rep.add(nmo + " scatter-add of dQ rows", acc.rows(S, uniq_nodes), ref_S)
        dIn_ref = np.asarray(A_R @ acc.rows(S, cols_R), dtype=F32)                                        # A_hat^T = A_hat
        dIn_name = nmo + " dH = A_hat.scatter(dQ)"
    else:
        dP = lo._operand("dP", lo._in.shape[0], lo.num_units)
        dZ = zbuf(lo)
        cs, wr = colsum_and_rows(acc.node_chunks([lo._in, dP]), wsel)
        rep.add(nmo + " db = colsum(dP)", grad_of(lo, "b"), cs)
        rep.add(nmo + " dZ = A_hat.dP", acc.rows(dZ, R), np.asarray(A_R @ acc.rows(dP, cols_R), dtype=F32))
        _, wr = colsum_and_rows(acc.node_chunks([lo._in, dZ]), wsel)
        rep.add(nmo + " dW rows = H^T.dZ", grad_of(lo, "W")[wsel], wr)
        dIn_ref = _dot64(acc.rows(dZ, R), W.T)
        dIn_name = nmo + " dH = dZ.W^T"
    li = layers.index(lo)
    prev = layers[li - 1]
    if type(prev) in (L.SparseConvolutionDenseLayer, L.ConvolutionDenseLayer) and prev.nonlinearity in ("rectify", "tanh"):
        dIn_ref = dIn_ref * _dact(prev.nonlinearity, acc.rows(prev._out, R))
        dIn_name += " * act'"
    grad_buf = lo._buf["dIn"]
    rep.add(dIn_name, acc.rows(grad_buf, R), dIn_ref)

    # ---- hidden layers, last to first
    for i in range(li - 1, -1, -1):
        ly = layers[i]
        nm = names[id(ly)]
        W = P0[(id(ly), "W")]
        prev = layers[i - 1] if i > 0 else None
        mask_prev = prev is not None and type(prev) in (L.SparseConvolutionDenseLayer, L.ConvolutionDenseLayer) \
            and prev.nonlinearity in ("rectify", "tanh")
        if isinstance(ly, L.HighwayConvolutionDenseLayer):
            dO_R = acc.rows(grad_buf, R)
            g_R, hc_R, in_R = acc.rows(ly._g, R), acc.rows(ly._Hc, R), acc.rows(ly._in, R)
            dP_ref = (g_R * dO_R * _dact(ly.nonlinearity, hc_R)).astype(F32)
            dG_ref = (dO_R * (hc_R - in_R) * g_R * (F32(1) - g_R)).astype(F32)
            dPb = ly._operand("dP", ly._in.shape[0], ly.num_units)
            dGb = ly._buf["dG"]
            rep.add(nm + " dP = g*dO*act'", acc.rows(dPb, R), dP_ref)
            rep.add(nm + " dGpre = dO*(H'-H)*g*(1-g)", acc.rows(dGb, R), dG_ref)
            dZb = zbuf(ly)
            rep.add(nm + " dZ = A_hat.dP", acc.rows(dZb, R), np.asarray(A_R @ acc.rows(dPb, cols_R), dtype=F32))
            wsel = np.sort(rng.choice(W.shape[0], size=min(n_param_rows, W.shape[0]), replace=False))
            cs_p, cs_g, w_z, w_g = None, None, None, None
            for a, p_, g_, z_ in acc.node_chunks([ly._in, dPb, dGb, dZb]):
                a64 = a[:, wsel].astype(np.float64).T
                cs_p = p_.astype(np.float64).sum(0) + (0 if cs_p is None else cs_p)
                cs_g = g_.astype(np.float64).sum(0) + (0 if cs_g is None else cs_g)
                w_z = a64 @ z_.astype(np.float64) + (0 if w_z is None else w_z)
                w_g = a64 @ g_.astype(np.float64) + (0 if w_g is None else w_g)
            rep.add(nm + " db = colsum(dP)", grad_of(ly, "b"), cs_p)
            rep.add(nm + " dbg = colsum(dGpre)", grad_of(ly, "bg"), cs_g)
            rep.add(nm + " dW rows = H^T.dZ", grad_of(ly, "W")[wsel], w_z)
            rep.add(nm + " dWg rows = H^T.dGpre", grad_of(ly, "Wg")[wsel], w_g)
            Wg = P0[(id(ly), "Wg")]
            dIn = ((1.0 - g_R.astype(np.float64)) * dO_R + _dot64(acc.rows(dZb, R), W.T) + _dot64(acc.rows(dGb, R), Wg.T))
            nmd = nm + " dH = (1-g)*dO + dZ.W^T + dGpre.Wg^T"
            if mask_prev:
                dIn = dIn * _dact(prev.nonlinearity, acc.rows(prev._out, R))
                nmd += " * act'"
            grad_buf = ly._buf["dIn"]
            rep.add(nmd, acc.rows(grad_buf, R), dIn)
            preact = mask_prev
        elif isinstance(ly, L.SparseConvolutionDenseLayer):
            dP1 = grad_buf                                   # already multiplied by act' (fused upstream)
            dZb = zbuf(ly)
            rep.add(nm + " dZ1 = A_hat.dP1", acc.rows(dZb, R), np.asarray(A_R @ acc.rows(dP1, cols_R), dtype=F32))
            V = X_host.shape[1]
            # vocabulary rows: the most frequent terms, the rarest, and random ones
            df = np.bincount(X_host.indices, minlength=V)
            byf = np.argsort(-df, kind="stable")
            vsel = np.unique(np.concatenate([byf[:max(2, n_param_rows // 4)], byf[-max(2, n_param_rows // 4):],
                                             rng.choice(V, size=n_param_rows, replace=False)]))
            Xc = sp.csc_matrix(X_host)[:, vsel]
            XTs = sp.csr_matrix(Xc.T).astype(np.float64)
            cs, w1 = None, np.zeros((len(vsel), ly.num_units), np.float64)
            s0 = 0
            for p_, z_ in acc.node_chunks([dP1, dZb]):
                cs = p_.astype(np.float64).sum(0) + (0 if cs is None else cs)
                w1 += XTs[:, s0:s0 + z_.shape[0]] @ z_.astype(np.float64)
                s0 += z_.shape[0]
            rep.add(nm + " db1 = colsum(dP1)", grad_of(ly, "b"), cs)
            rep.add(nm + " dW1 rows = X^T.dZ1", grad_of(ly, "W")[vsel], w1)
        else:
            raise NotImplementedError("sampled parity: hidden ConvolutionDenseLayer without gate")

    # ---- Adam + elastic net (mlpconv.py:235-245, :263), float32 as Lasagne computes it
    lr, b1, b2, eps = (F32(x) for x in adam.hyper)
    one = F32(1)
    t_new = F32(t_before[0] + one)
    a_t = F32(lr * np.sqrt(one - b2 ** t_new, dtype=F32) / (one - b1 ** t_new))
    for ly in layers:
        for pname, t, tags in ly.params:
            key = (id(ly), pname)
            sel = prow[key]
            p0 = P0[key] if sel is None else P0[key][sel]
            g = host(ly.grads[pname]) if sel is None else host(ly.grads[pname])[sel]
            coef = F32(adam._reg[idx_of[t.data_ptr()]])
            g = (g + F32(0.5) * coef * (np.sign(p0) + F32(2) * p0)).astype(F32)
            m0, v0 = st0[key]
            m1 = b1 * m0 + (one - b1) * g
            v1 = b2 * v0 + (one - b2) * g * g
            ref = (p0 - a_t * m1 / (np.sqrt(v1, dtype=F32) + eps)).astype(F32)
            got = host(t) if sel is None else host(t)[sel]
            rep.add("%s adam %s" % (names[id(ly)], pname), got, ref)
    return rep.summary()
