"""Worker for the multi-process tests: run under torchrun (NCCL, one rank per GPU) or spawned with
gloo on CPU.  Compares the row-partitioned model with the single-process oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def gpu_main():
    import torch
    import torch.distributed as dist
    from graphconvgeo_b200 import synth
    from graphconvgeo_b200.dist import DistMLPCONV
    from oracle import gcn_oracle as go
    from util import assert_close
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import datetime
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    w = synth.make_workload("tiny")
    # the last case runs with drop_out=True and p = 0: the DropoutLayer sits in the chain (its branch of the
    # backward must hand the layer below a gradient w.r.t. the ACTIVATION, so relu' is applied there) while the
    # masks are the identity, so the oracle without dropout is the reference
    cases = [
            (2, False, None, "row", False, False), (3, True, "labels", "row", False, False),
            (2, False, "labels", "feature", False, False), (3, True, None, "feature", False, False),
            (3, True, "labels", "feature", True, False), (2, False, None, "feature", True, False),
            (3, True, None, "row", False, True), (2, False, "labels", "feature", True, True),
            # Twitter-scale code paths forced on the small input: dense head of X on the tensor-core engine, document-
            # blocked X^T.dZ1, dW1 summed over ranks piece by piece (the pieces must mean the same on every rank)
            (3, True, "labels", "feature", True, "big"), (2, False, None, "row", False, "big"),
            # the collectives issued from libgcg.so itself (gcg_comm_init, gcg_spmm_rowpart_allgather_f32,
            # gcg_allreduce_grads_f32: the multi-GPU C ABI of SURVEY section 8b) instead of torch.distributed
            (3, True, "labels", "row", False, False, "native"), (2, False, None, "row", False, "big", "native"),
            (3, True, None, "feature", True, False, "native")]
    for case in cases:
        n_layers, highway, reorder, partition, peer, drop = case[:6]
        collectives = case[6] if len(case) > 6 else "torch"
        big = drop == "big"
        drop = drop is True
        os.environ["GCG_X_FORCE_BIG"] = "1" if big else "0"
        os.environ["GCG_X_HEAD"] = "32" if big else "256"
        os.environ["GCG_XT_BLOCK_MB"] = "0.015" if big else "96"
        os.environ["GCG_XT_BLOCK_MIN_COLS"] = "128" if big else "1024"
        os.environ["GCG_XT_HEAVY_FACTOR"] = "1" if big else "16"
        from graphconvgeo_b200 import ops as _ops
        _ops.set_gemm_mode("tf32x3" if big else "auto")
        rng = np.random.RandomState(5)
        params = go.init_params(rng, w.X.shape[1], w.hidden, w.n_classes, n_layers, highway)
        idx = rng.choice(w.train_indices, size=len(w.train_indices)).astype(np.int32)      # duplicates
        y = w.Y[idx].astype(np.int32)
        net = go.GCNOracle(w.X, w.A_hat, n_layers, highway, (1e-4, 2e-4))
        ref_params = [p.copy() for p in params]
        hist = go.train_epochs(net, ref_params, idx, y, 3)
        m = DistMLPCONV(n_epochs=1, regul_coefs=[1e-4, 2e-4], hidden_layer_size=w.hidden, n_layers=n_layers,
                        highway=highway, init_parameters=[p.copy() for p in params], device=dev, reorder=reorder,
                        partition=partition, peer_memory=peer, drop_out=drop, dropout_coefs=[0.0, 0.0],
                        collectives=collectives)
        m.prepare(w.X, idx, w.dev_indices, w.test_indices, w.Y, w.A_hat)
        assert m.part.world == world and (m.part.comm is not None) == (collectives == "native")
        for step in range(3):
            m.f_train()
            l, a = m.train_results()
            assert abs(l - hist[step][0]) <= 2e-5 * abs(hist[step][0]), (step, l, hist[step])
            assert abs(a - hist[step][1]) <= 2.0 / len(idx)
            if step == 0:       # activations of the first step, gathered over ranks, original node order
                c = net.forward(params, idx)
                conv_layers = [ly for ly in m.layers[:-1] if hasattr(ly, "W")]
                for i, ly in enumerate(conv_layers):
                    assert_close(m.node_rows(ly._out), c["A"][i], what="activation %d" % i)
        for p_gpu, p in zip(m.get_param_values(), ref_params):
            assert_close(p_gpu, p, atol=1e-5, rtol=1e-3, what="params after 3 steps")
        if big:
            l1 = m.l_hid1
            assert getattr(l1, "_x_head", None) is not None and getattr(l1, "_xt_blocked", None) is not None
            assert len(l1._xt_blocked[1].blocks) > 1
        from graphconvgeo_b200 import lasagne_layers as L
        L.set_all_param_values(m.l_out, ref_params)
        proba = m.predict_proba("test")
        assert_close(proba, net.predict_proba(ref_params, w.test_indices), atol=1e-6)
        pred = m.predict("dev")
        refp = net.predict_proba(ref_params, w.dev_indices)
        top2 = np.sort(refp, axis=1)[:, -2:]
        tie = (top2[:, 1] - top2[:, 0]) <= 1e-6
        assert np.array_equal(pred[~tie], refp.argmax(-1)[~tie])
        acc = m.accuracy("test", w.Y[w.test_indices].astype(np.int32))
        _, ref_acc = net.loss_acc(ref_params, w.test_indices, w.Y[w.test_indices].astype(np.int32))
        assert abs(acc - ref_acc) <= 2.0 / len(w.test_indices)
        if rank == 0:
            print("dist case", n_layers, highway, reorder, partition, "peer" if (peer and m.part.peer is not None) else "nccl",
                  "dropout(p=0)" if drop else "", "forced Twitter-scale paths" if big else "",
                  "collectives: " + collectives, "OK", flush=True)
        if m.part.peer is not None:
            m.part.peer.check()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_GPU_OK", flush=True)


def cpu_main(rank, world, port):
    """gloo, CPU: the host-side logic of the row partition (no kernels): block split, in-place
    all-gather bookkeeping, target-index ownership; SpMM emulated with scipy."""
    import scipy.sparse as sp
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from graphconvgeo_b200 import synth
    from graphconvgeo_b200.dist import DistCSRMatrix, DistTargetIndices, RowPartition
    from graphconvgeo_b200.sparse import CSRMatrix
    w = synth.make_workload("tiny")
    A = w.A_hat
    n = A.shape[0]
    part = RowPartition(n, world, rank, torch.device("cpu"))
    assert part.n_pad >= n and part.n_loc % 4 == 0
    Ag = CSRMatrix.from_scipy(A, device="cpu")
    Hd = DistCSRMatrix.from_global(Ag, part)
    to_sp = lambda m: sp.csr_matrix((m.host[2], m.host[1], m.host[0]), shape=m.shape)
    diag, off = to_sp(Hd.diag), to_sp(Hd.off)
    ref_rows = sp.csr_matrix(A[part.r0:part.r1])
    got = (diag + off)[:part.r1 - part.r0, :n]
    assert (got != ref_rows).nnz == 0
    c0 = part.rank * part.n_loc
    assert diag.nnz == 0 or (diag.indices.min() >= c0 and diag.indices.max() < c0 + part.n_loc)
    assert off.nnz == 0 or np.all((off.indices < c0) | (off.indices >= c0 + part.n_loc))
    # in-place all-gather of a dense operand, then the two-phase product == rows of A @ Z
    rng = np.random.RandomState(0)
    Z = rng.standard_normal((n, 12)).astype(np.float32)
    full = part.full("Z", 12)
    loc = part.local_view(full)
    loc[:part.r1 - part.r0] = torch.from_numpy(Z[part.r0:part.r1])
    part.all_gather_async(full).wait()
    assert np.array_equal(full[:n].numpy(), Z)
    out = diag @ full.numpy() + off @ full.numpy()
    assert np.allclose(out[:part.r1 - part.r0], A[part.r0:part.r1] @ Z, atol=1e-5)
    # target indices: every position owned exactly once across ranks; sub-matrix rows match
    idx = rng.choice(n, size=333).astype(np.int32)
    ti = DistTargetIndices(idx, Hd, len(idx))
    cnt = torch.zeros(len(idx))
    cnt[torch.from_numpy(ti.sel)] = 1
    dist.all_reduce(cnt)
    assert torch.all(cnt == 1)
    sub = ti.Hsub
    sub_sp = to_sp(sub.diag) + to_sp(sub.off)
    assert (sub_sp[:, :n] != sp.csr_matrix(A[idx[ti.sel]])).nnz == 0
    ptr, pos = None, None
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    gpu_main()
