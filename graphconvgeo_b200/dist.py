"""Row-partitioned multi-GPU execution of the GCN hot path (one process per GPU).

The reference is single-process; this is the scaling design BASELINE.json's
north_star asks for:
  * rank p owns the contiguous node range [p*n_loc, (p+1)*n_loc) of A_hat (all columns), of X,
    of every activation / gradient slab and of the target indices falling in its range;
    parameters and Adam state are replicated.
  * every propagation A_hat.Z needs all rows of the dense operand: the local slab lives inside a
    full [P*n_loc, F] buffer and is all-gathered IN PLACE with NCCL over NVLink (asynchronously,
    on NCCL's stream) while the main stream already runs the SpMM over the DIAGONAL column block
    (local columns only read the local slab); the off-diagonal block follows with
    accumulate + the fused epilogue once the gather has landed.
  * parameter gradients are partial sums over local rows -> all-reduce (async, overlapped with
    the rest of the backward); loss / accuracy partial sums -> one small all-reduce.
  * X.W1, all dense GEMMs, epilogues, the gate and Adam need no communication.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import lasagne_layers as L
from . import ops
from .mlpconv import MLPCONV
from .sparse import CSRMatrix, as_csr


class NativeComm:
    """gcg_comm (include/gcg.h): the collectives of the row-partitioned epoch issued from libgcg.so itself -- NCCL
    bound at run time, its own communication stream, event fences instead of Work handles.  torch.distributed is
    only used to hand rank 0's NCCL unique id to the other ranks."""

    class _Fence:
        def __init__(self, comm):
            self.comm = comm

        def wait(self):
            self.comm.wait()

    def __init__(self, group=None):
        import ctypes as C
        from . import _lib
        Lb = _lib.lib()
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        uid = C.create_string_buffer(128)
        if self.rank == 0:
            _lib.check(Lb.gcg_comm_unique_id(uid), "gcg_comm_unique_id")
        box = [bytes(uid.raw)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self._uid = C.create_string_buffer(box[0], 128)
        h = C.c_void_p()
        _lib.check(Lb.gcg_comm_init(self._uid, self.world, self.rank, C.byref(h)), "gcg_comm_init")
        self._h = h

    def wait(self):
        from . import _lib
        _lib.check(_lib.lib().gcg_comm_wait(self._h, ops._stream()), "gcg_comm_wait")

    def all_reduce(self, tensors, wait=False):
        """in-place sum over ranks of contiguous float32 tensors (one NCCL group); returns a fence with wait()"""
        import ctypes as C
        from . import _lib
        ts = [t for t in (tensors if isinstance(tensors, (list, tuple)) else [tensors])]
        for t in ts:
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise TypeError("NativeComm.all_reduce needs contiguous float32 CUDA tensors")
        n = len(ts)
        ptrs = (C.c_void_p * max(n, 1))(*[t.data_ptr() for t in ts])
        sizes = (C.c_int64 * max(n, 1))(*[t.numel() for t in ts])
        _lib.check(_lib.lib().gcg_allreduce_grads_f32(self._h, n, ptrs, sizes, int(bool(wait)), ops._stream()),
                   "gcg_allreduce_grads_f32")
        return NativeComm._Fence(self)

    def all_gather(self, full, floats_per_rank, wait=True):
        """in-place all-gather of ``full`` = [world][floats_per_rank] (gcg_allgather_rows_f32)"""
        import ctypes as C
        from . import _lib
        _lib.check(_lib.lib().gcg_allgather_rows_f32(self._h, C.c_void_p(full.data_ptr()), int(floats_per_rank),
                                                     int(bool(wait)), ops._stream()), "gcg_allgather_rows_f32")
        return NativeComm._Fence(self)

    def spmm_rowpart(self, diag, off, full, F, n_loc, out, **epi):
        """gcg_spmm_rowpart_allgather_f32: in-place all-gather of ``full`` ([world*n_loc, ld] buffer, F <= ld valid
        columns) overlapped with the diagonal block"""
        from . import _lib
        fp, ld = ops._mat(full, "Z_full")
        cp, ldc = ops._mat(out, "out")
        gate, carry, conv = epi.get("gate"), epi.get("carry"), epi.get("conv_out")
        gp = hp = vp = None
        ldg = ldh = ldv = 0
        if gate is not None:
            gp, ldg = ops._mat(gate, "gate")
            hp, ldh = ops._mat(carry, "carry")
            if conv is not None:
                vp, ldv = ops._mat(conv, "conv_out")
        ws, wsb = ops.scratch.get(max(diag.workspace_bytes(ldc), off.workspace_bytes(ldc)), full.device)
        pc = ops.auto_panel_cols(off.shape[1], F, off.nnz, getattr(off, "spmm_mode", None))
        _lib.check(_lib.lib().gcg_spmm_rowpart_allgather_f32(
            self._h, diag.plan, off.plan, fp, ld, F, int(n_loc), cp, ldc, ops._vec(epi.get("bias"), "bias"),
            _lib.act_code(epi.get("act", "identity")), gp, ldg, hp, ldh, vp, ldv, int(pc), ws, wsb, ops._stream()),
            "gcg_spmm_rowpart_allgather_f32")
        return out

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                from . import _lib
                _lib.lib().gcg_comm_destroy(self._h)
                self._h = None
        except Exception:
            pass


class RowPartition:
    """Equal contiguous row blocks (padded): n_loc = ceil(N / P) rounded up to 4 rows."""

    def __init__(self, n_total, world, rank, device=None, group=None):
        self.n_total, self.world, self.rank = int(n_total), int(world), int(rank)
        self.n_loc = (-(-self.n_total // self.world) + 3) // 4 * 4
        self.n_pad = self.n_loc * self.world
        self.r0 = min(self.n_total, self.rank * self.n_loc)
        self.r1 = min(self.n_total, self.r0 + self.n_loc)
        self.device = device
        self.group = group
        self._full = {}
        self.bytes_gathered = 0
        self.comm = None            # NativeComm: collectives through the C ABI instead of torch.distributed

    def owner(self, rows):
        return np.asarray(rows) // self.n_loc

    def local_rows(self, rows):
        """positions (into ``rows``) owned by this rank, and their local row ids"""
        rows = np.asarray(rows, dtype=np.int64)
        sel = np.flatnonzero((rows >= self.r0) & (rows < self.r1))
        return sel, (rows[sel] - self.rank * self.n_loc).astype(np.int32)

    # full [n_pad, F] operand buffers, one per (key, F); the local slab is a view into it
    def full(self, key, F):
        k = (key, int(F))
        buf = self._full.get(k)
        if buf is None:
            buf = ops.alloc_mat(self.n_pad, F, self.device, zero=True)
            self._full[k] = buf
        return buf

    def local_view(self, full):
        return full[self.rank * self.n_loc:(self.rank + 1) * self.n_loc]

    def all_gather_async(self, full):
        """in-place NCCL all-gather of the slabs of ``full``; returns the Work handle."""
        base = full._base if full._base is not None else full
        flat = base.view(-1)
        per = flat.numel() // self.world
        self.bytes_gathered += (self.world - 1) * per * 4
        return dist.all_gather_into_tensor(flat, flat[self.rank * per:(self.rank + 1) * per],
                                           group=self.group, async_op=True)


def split_columns(host, c0, c1):
    """(indptr, indices, vals) -> (diag part with columns in [c0, c1), the rest); order preserved."""
    ip, ix, d = host
    n = len(ip) - 1
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(ip))
    m = (ix >= c0) & (ix < c1)

    def take(mask):
        cnt = np.bincount(rows[mask], minlength=n)
        p = np.zeros(n + 1, np.int32)
        np.cumsum(cnt, out=p[1:])
        return (p, np.ascontiguousarray(ix[mask]), np.ascontiguousarray(d[mask]))

    return take(m), take(~m)


class DistCSRMatrix:
    """Rows owned by this rank of a row-partitioned sparse matrix whose dense operand is
    row-partitioned too.  ``spmm`` = async all-gather of the operand overlapped with the
    diagonal-block SpMM, then the off-diagonal block with accumulate + epilogue."""

    def __init__(self, part: RowPartition, host, n_rows, long_row_threshold=256):
        self.part = part
        self.host = host
        self.shape = (int(n_rows), part.n_pad)
        self.long_row_threshold = long_row_threshold
        c0 = part.rank * part.n_loc
        dg, off = split_columns(host, c0, c0 + part.n_loc)
        self.diag = CSRMatrix.from_host(dg, self.shape, part.device, long_row_threshold)
        self.off = CSRMatrix.from_host(off, self.shape, part.device, long_row_threshold)
        self.nnz = len(host[1])
        self.diag_fraction = (len(dg[1]) / max(1, self.nnz))

    @property
    def device(self):
        return self.part.device

    @classmethod
    def from_global(cls, A: CSRMatrix, part: RowPartition):
        ip, ix, d = A._host_arrays()
        lo, hi = part.r0, part.r1
        p = np.zeros(part.n_loc + 1, np.int32)
        p[:hi - lo + 1] = ip[lo:hi + 1] - ip[lo]
        p[hi - lo + 1:] = p[hi - lo]
        host = (p, np.ascontiguousarray(ix[ip[lo]:ip[hi]]), np.ascontiguousarray(d[ip[lo]:ip[hi]]))
        return cls(part, host, part.n_loc, A.long_row_threshold)

    def gather_rows(self, local_idx):
        local_idx = np.asarray(local_idx, dtype=np.int64)
        ip, ix, d = self.host
        lens = (ip[local_idx + 1] - ip[local_idx]).astype(np.int64)
        p = np.zeros(len(local_idx) + 1, np.int32)
        np.cumsum(lens, out=p[1:])
        starts = ip[local_idx].astype(np.int64)
        take = np.repeat(starts - p[:-1].astype(np.int64), lens) + np.arange(int(p[-1]), dtype=np.int64)
        host = (p, np.ascontiguousarray(ix[take]), np.ascontiguousarray(d[take]))
        return DistCSRMatrix(self.part, host, len(local_idx), self.long_row_threshold)

    def operand(self, key, F):
        """local slab [n_loc, F] that lives inside the full gather buffer (no staging copy)."""
        return self.part.local_view(self.part.full(key, F))

    def dist_spmm(self, B, out=None, key="stage", **epi):
        part = self.part
        F = B.shape[1]
        full = None
        for (k, f), buf in part._full.items():           # is B already a slab of a gather buffer?
            if f == F and part.local_view(buf).data_ptr() == B.data_ptr():
                full = buf
                break
        if full is None:
            full = part.full(key, F)
            part.local_view(full).copy_(B)
        if out is None:
            out = ops.alloc_mat(self.shape[0], F, B.device)
        base = full._base if full._base is not None else full
        if part.comm is not None:
            assert base.is_contiguous() and base.shape[0] == part.n_pad
            part.bytes_gathered += (part.world - 1) * part.n_loc * base.stride(0) * 4
            if self.shape[0] == 0:      # no rows here (a rank without targets): still a party to the collective
                part.comm.all_gather(base, part.n_loc * base.stride(0))
                return out
            # the whole propagation (gather, diagonal block, fence, off-diagonal block + epilogue) is ONE C call
            return part.comm.spmm_rowpart(self.diag, self.off, base, F, part.n_loc, out, **epi)
        work = part.all_gather_async(full)                # NCCL stream; waits for what is enqueued so far
        if self.shape[0] > 0:
            ops.spmm(self.diag, full, out=out)            # local columns: only the local slab is read
        work.wait()                                       # main stream waits for the gather
        if self.shape[0] == 0:
            return out
        return ops.spmm(self.off, full, out=out, accumulate=True, **epi)


class _PhaseProfile:
    """GCG_DIST_PROFILE=1: CUDA-event timing of the phases of the feature-sliced propagation (rank 0 prints)."""

    def __init__(self):
        import os
        self.on = os.environ.get("GCG_DIST_PROFILE") == "1"
        self.acc = {}
        self.pending = []

    def mark(self, name):
        if not self.on:
            return
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.pending.append((name, e))

    def flush(self):
        if not self.on or len(self.pending) < 2:
            self.pending = []
            return
        torch.cuda.synchronize()
        for (n0, e0), (n1, e1) in zip(self.pending[:-1], self.pending[1:]):
            if n1 == "start":
                continue
            a = self.acc.setdefault(n1, [0, 0.0])
            a[0] += 1
            a[1] += e0.elapsed_time(e1)
        self.pending = []

    def report(self):
        self.flush()
        return {k: {"calls": c, "ms": t} for k, (c, t) in self.acc.items()}


phase_profile = _PhaseProfile()

import os as _os
# feature-sliced propagation: second transpose fused into the SpMM epilogue (0 = separate gcg_push_rows_f32 pass)
_FUSED_PUSH = _os.environ.get("GCG_DIST_FUSED_PUSH", "1") != "0"


class _RawCudaArray:
    def __init__(self, ptr, numel):
        self.__cuda_array_interface__ = {"shape": (int(numel),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerBuffers:
    """IPC-shared staging buffers for the peer-memory transposes (gcg_push_cols_f32 / gcg_push_rows_f32):
    every rank owns a `recv` buffer ([n_pad, Fp] column slice of all nodes) and a `back` buffer
    ([P][rows][Fp] row groups coming home) that the PEERS write with plain stores over NVLink."""

    def __init__(self, part: RowPartition, fp_max, rows_max):
        import ctypes as C
        from . import _lib
        L = _lib.lib()
        self.part = part
        P = part.world
        self.recv_floats = int(part.n_pad * fp_max)
        self.back_floats = int(P * rows_max * fp_max)
        self._own, handles = [], []
        for nfl in (self.recv_floats, self.back_floats, 64):       # the third block: 64 int32 barrier flags (zeroed)
            ptr = C.c_void_p()
            h = C.create_string_buffer(64)
            _lib.check(L.gcg_peer_alloc(nfl * 4, C.byref(ptr), h), "gcg_peer_alloc")
            self._own.append(ptr.value)
            handles.append(h.raw)
        torch.cuda.synchronize(part.device)        # the zero fill has landed before any peer can write a flag
        allh = [None] * P
        dist.all_gather_object(allh, handles, group=part.group)
        self.recv_ptrs, self.back_ptrs, self.flag_ptrs, self._opened = [], [], [], []
        for q in range(P):
            if q == part.rank:
                self.recv_ptrs.append(self._own[0])
                self.back_ptrs.append(self._own[1])
                self.flag_ptrs.append(self._own[2])
                continue
            got = []
            for h in allh[q]:
                ptr = C.c_void_p()
                _lib.check(L.gcg_peer_open(C.create_string_buffer(h, 64), C.byref(ptr)), "gcg_peer_open")
                got.append(ptr.value)
                self._opened.append(ptr.value)
            self.recv_ptrs.append(got[0])
            self.back_ptrs.append(got[1])
            self.flag_ptrs.append(got[2])
        dev = part.device
        self.recv_t = torch.as_tensor(_RawCudaArray(self._own[0], self.recv_floats), device=dev)
        self.back_t = torch.as_tensor(_RawCudaArray(self._own[1], self.back_floats), device=dev)
        self._flag = torch.zeros(1, dtype=torch.float32, device=dev)
        self._err = torch.zeros(1, dtype=torch.int32, device=dev)
        self._seq = 0
        import os
        self.nccl_barrier = os.environ.get("GCG_DIST_NCCL_BARRIER") == "1"
        self._C = C
        self._L = L
        dist.barrier(group=part.group)             # every rank has mapped every buffer

    def barrier(self):
        """cross-rank ordering point in stream order, no host synchronisation: peer-written flags waited on by
        a one-warp kernel (gcg_peer_barrier); GCG_DIST_NCCL_BARRIER=1 selects the tiny NCCL all-reduce of round 1"""
        if self.nccl_barrier:
            dist.all_reduce(self._flag, group=self.part.group)
            return
        from . import _lib
        C = self._C
        self._seq += 1
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(self._L.gcg_peer_barrier(self.ptr_array(self.flag_ptrs), C.c_void_p(self._own[2]), self.part.world,
                                            self.part.rank, self._seq, C.c_void_p(self._err.data_ptr()), stream),
                   "gcg_peer_barrier")

    def check(self):
        """raise if a barrier gave up waiting for a peer (call at a host synchronisation point)"""
        if int(self._err.item()) != 0:
            raise RuntimeError("gcg_peer_barrier: a peer GPU did not arrive within the time limit")

    def ptr_array(self, ptrs, byte_offsets=None):
        C = self._C
        vals = [p + (0 if byte_offsets is None else byte_offsets[i]) for i, p in enumerate(ptrs)]
        return (C.c_void_p * len(vals))(*vals)


class FeatureSplitCSRMatrix:
    """Alternative multi-GPU scheme for A_hat.Z (SURVEY 7.2 "feature-partitioned SpMM"): every rank keeps
    the WHOLE sparse matrix (A_hat is only 8*nnz bytes) and propagates a COLUMN slice Z[:, cols_p] of the
    dense operand for all nodes, so the SpMM itself needs no communication.  The operand arrives row-
    partitioned from the local GEMM, so it is transposed with one all-to-all before ([n_loc, F] ->
    [N, F/P]) and one after the SpMM: each rank sends n_loc*F*(P-1)/P floats per direction -- 8x less than
    the all-gather of the row-partitioned scheme at P = 8 (measured: DESIGN.md section 5).
    Output rows are grouped by owner rank (``row_counts``); bias + activation stay fused in the SpMM
    (they are column-wise), the highway mix runs after the transpose back."""

    def __init__(self, part: RowPartition, full: CSRMatrix, row_counts):
        self.part = part
        self.full = full
        self.row_counts = [int(c) for c in row_counts]
        self.shape = (self.row_counts[part.rank], part.n_pad)
        self.long_row_threshold = full.long_row_threshold
        self.nnz = full.nnz
        self.diag_fraction = None
        self.host = full.host

    @property
    def device(self):
        return self.part.device

    @classmethod
    def from_global(cls, A: CSRMatrix, part: RowPartition):
        ip, ix, d = A._host_arrays()
        p = np.empty(part.n_pad + 1, np.int32)
        p[:len(ip)] = ip
        p[len(ip):] = ip[-1]
        full = CSRMatrix.from_host((p, ix, d), (part.n_pad, part.n_pad), part.device, A.long_row_threshold)
        return cls(part, full, [part.n_loc] * part.world)

    def gather_rows_grouped(self, idx_global):
        """A[idx,:] for ALL target nodes, rows grouped by owner rank (ascending position inside a group)."""
        idx_global = np.asarray(idx_global, dtype=np.int64)
        owners = self.part.owner(idx_global)
        order = np.argsort(owners, kind="stable")
        counts = np.bincount(owners, minlength=self.part.world)
        sub = self.full.gather_rows(idx_global[order].astype(np.int32))
        sub.shape = (sub.shape[0], self.part.n_pad)
        return FeatureSplitCSRMatrix(self.part, sub, counts)

    def _buf(self, key, shape):
        k = ("fs", key, tuple(shape))
        b = self.part._full.get(k)
        if b is None:
            b = torch.zeros(shape, dtype=torch.float32, device=self.part.device)
            self.part._full[k] = b
        return b

    def dist_spmm(self, B, out=None, bias=None, act="identity", gate=None, carry=None, conv_out=None):
        part = self.part
        P, r, n_loc = part.world, part.rank, part.n_loc
        F = B.shape[1]
        assert B.shape[0] == n_loc
        Fp = (-(-F // P) + 3) // 4 * 4
        my_rows = self.row_counts[r]
        if out is None:
            out = ops.alloc_mat(my_rows, F, B.device)
        peer = getattr(part, "peer", None)
        rows_total = int(sum(self.row_counts))
        if peer is not None and part.n_pad * Fp <= peer.recv_floats and P * my_rows * Fp <= peer.back_floats \
                and all(P * c * Fp <= peer.back_floats for c in self.row_counts):
            return self._dist_spmm_peer(peer, B, out, bias, act, gate, carry, conv_out, Fp, rows_total, my_rows)
        # 1. row layout -> column slices (zero padded), one all-to-all
        send = self._buf("send", (P, n_loc, Fp))
        ops.pack_cols(B, P, Fp, send)
        recv = self._buf("recv", (P * n_loc, Fp))
        dist.all_to_all_single(recv.view(-1), send.view(-1), group=part.group)
        part.bytes_gathered += (P - 1) * n_loc * Fp * 4
        # 2. local SpMM over ALL rows for this rank's columns; bias/act are column-wise -> fused
        bs = None
        if bias is not None:
            bpad = self._buf("bias", (P * Fp,))
            bpad[:F].copy_(bias)
            bs = bpad[r * Fp:(r + 1) * Fp]
        outslice = self._buf("outslice", (max(rows_total, 1), Fp))
        if rows_total > 0:
            ops.spmm(self.full, recv, out=outslice[:rows_total], bias=bs, act=act)
        # 3. column slices -> row layout (rows grouped by owner), second all-to-all
        back = self._buf("back", (P * max(my_rows, 1), Fp))
        dist.all_to_all_single(back[:P * my_rows], outslice[:rows_total], output_split_sizes=[my_rows] * P,
                               input_split_sizes=self.row_counts, group=part.group)
        if my_rows == 0:
            return out
        dst = out if gate is None or conv_out is None else conv_out
        ops.unpack_cols(back[:P * my_rows], P, Fp, dst)
        if gate is not None:
            ops.highway_mix(dst, gate, carry, out=out)       # out = g*Hc + (1-g)*H ; Hc kept in conv_out
        return out

    def _dist_spmm_peer(self, peer, B, out, bias, act, gate, carry, conv_out, Fp, rows_total, my_rows):
        """Same propagation with the two transposes done by peer-memory stores (no NCCL all-to-all)."""
        import ctypes as C
        from . import _lib
        L = _lib.lib()
        part = self.part
        P, r, n_loc = part.world, part.rank, part.n_loc
        F = B.shape[1]
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        bp, ldb = ops._mat(B, "B")
        pp = phase_profile
        pp.mark("start")
        # 1. every rank writes slice q of its rows into rank q's recv buffer (rows [r*n_loc, (r+1)*n_loc))
        _lib.check(L.gcg_push_cols_f32(bp, ldb, n_loc, F, P, Fp, peer.ptr_array(peer.recv_ptrs), r * n_loc, stream),
                   "gcg_push_cols_f32")
        part.bytes_gathered += (P - 1) * n_loc * Fp * 4
        pp.mark("push_cols")
        peer.barrier()
        pp.mark("barrier1")
        recv = peer.recv_t[:part.n_pad * Fp].view(part.n_pad, Fp)
        bs = None
        if bias is not None:
            bpad = self._buf("bias", (P * Fp,))
            bpad[:F].copy_(bias)
            bs = bpad[r * Fp:(r + 1) * Fp]
        # 2. the row group of owner q goes to slot r of rank q's back buffer ([P][row_counts[q]][Fp])
        off = np.zeros(P + 1, np.int64)
        np.cumsum(self.row_counts, out=off[1:])
        offs = (C.c_int64 * (P + 1))(*[int(x) for x in off])
        dst_ptrs = peer.ptr_array(peer.back_ptrs, [r * self.row_counts[q] * Fp * 4 for q in range(P)])
        if _FUSED_PUSH:
            # ... written by the SpMM's own epilogue stores (gcg_spmm_csr_routed_f32): the transfer over NVLink
            # overlaps the gather instead of following it as a separate pass over the slice
            if rows_total > 0:
                A = self.full
                rp, ldr = ops._mat(recv, "recv")
                ws, wsb = ops.scratch.get(A.workspace_bytes(Fp), recv.device)
                _lib.check(L.gcg_spmm_csr_routed_f32(A.plan, rp, ldr, Fp, P, dst_ptrs, offs, Fp,
                                                     ops._vec(bs, "bias") if bs is not None else None,
                                                     _lib.act_code(act), 0, ws, wsb, stream), "gcg_spmm_csr_routed_f32")
            pp.mark("spmm+push_rows")
        else:
            outslice = self._buf("outslice", (max(rows_total, 1), Fp))
            if rows_total > 0:
                ops.spmm(self.full, recv, out=outslice[:rows_total], bias=bs, act=act)
            pp.mark("spmm")
            _lib.check(L.gcg_push_rows_f32(C.c_void_p(outslice.data_ptr()), offs, P, Fp, dst_ptrs, 0, stream),
                       "gcg_push_rows_f32")
            pp.mark("push_rows")
        peer.barrier()
        pp.mark("barrier2")
        if my_rows == 0:
            return out
        dst = out if gate is None or conv_out is None else conv_out
        ops.unpack_cols(peer.back_t[:P * my_rows * Fp], P, Fp, dst)
        if gate is not None:
            ops.highway_mix(dst, gate, carry, out=out)
        pp.mark("unpack")
        return out


class DistTargetIndices(L.TargetIndices):
    """The part of a target_indices vector whose nodes this rank owns."""

    def __init__(self, idx_global_perm, H: DistCSRMatrix, n_global):   # noqa: super().__init__ not wanted
        part = H.part
        self.sel, self.local = part.local_rows(idx_global_perm)
        self.idx_global = np.asarray(idx_global_perm)
        self.n = len(self.local)
        self.n_global = int(n_global)
        self.device = part.device
        self.H = H
        self.dev = torch.from_numpy(self.local).to(self.device)
        self._Hsub = None
        self._pos = None

    @property
    def Hsub(self):
        if self._Hsub is None:
            if hasattr(self.H, "gather_rows_grouped"):
                self._Hsub = self.H.gather_rows_grouped(self.idx_global)
            else:
                self._Hsub = self.H.gather_rows(self.local)
        return self._Hsub

    @property
    def positions(self):
        if self._pos is None:
            self._pos = ops.scatter_positions(self.local, self.H.shape[0], self.device)
        return self._pos


class DistMLPCONV(MLPCONV):
    """MLPCONV over a row-partitioned graph.  Every rank is given the same full inputs (as the
    single-process fit() is); each keeps its row block.  Results (loss, acc, predictions gathered
    over ranks, parameters) equal the single-GPU ones up to summation order of the all-reduces."""

    def __init__(self, *args, group=None, partition="auto", peer_memory=True, collectives=None, **kwargs):
        kwargs["cuda_graph"] = False          # NCCL work is enqueued eagerly
        kwargs["native_epoch"] = False        # collectives and peer barriers are not part of a gcg_epoch
        super().__init__(*args, **kwargs)
        assert partition in ("row", "feature", "auto")
        if partition == "auto":               # measured on B200: the all-gather wins at 2 ranks, the transposes from 4 up
            partition = "feature" if dist.get_world_size(group) >= 4 else "row"
        self.partition = partition            # how A_hat.Z is distributed (see the two matrix classes)
        self.peer_memory = peer_memory        # feature mode: transposes by P2P stores instead of NCCL all-to-all
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        # "torch": torch.distributed issues the collectives; "native": libgcg.so does (gcg_comm_*, gcg_allreduce_grads_f32,
        # gcg_spmm_rowpart_allgather_f32 -- the survey's multi-GPU C ABI).  Same NCCL underneath.
        import os
        self.collectives = collectives or os.environ.get("GCG_DIST_COMM", "torch")
        assert self.collectives in ("torch", "native")

    def prepare(self, X, train_indices, dev_indices, test_indices, Y, H):
        Y = np.asarray(Y)
        in_size = X.shape[1]
        out_size = int(np.max(Y)) + 1
        self.X, self.H = X, H
        self.train_indices, self.dev_indices, self.test_indices = train_indices, dev_indices, test_indices
        Xg = as_csr(X, self.device, long_row_threshold=1024)
        Hg = as_csr(H, self.device)
        n = Hg.shape[0]
        mode = self.reorder
        if isinstance(mode, str) and mode == "auto":
            mode = "labels"
        self.node_order = None
        inv = None
        if mode is not None:
            if isinstance(mode, str) and mode == "labels":
                order = np.argsort(Y[:n], kind="stable").astype(np.int32)
            elif isinstance(mode, str) and mode == "degree":
                order = np.argsort(-np.diff(Hg._host_arrays()[0]), kind="stable").astype(np.int32)
            else:
                order = np.ascontiguousarray(np.asarray(mode), dtype=np.int32)
            inv = np.empty(n, np.int32)
            inv[order] = np.arange(n, dtype=np.int32)
            Xg = Xg.permute(order)
            Hg = Hg.permute(order, col_map=inv)
            self.node_order, self.node_inverse = order, inv
        node_map = (lambda i: np.asarray(i)) if inv is None else (lambda i: inv[np.asarray(i)])
        self._node_map = node_map
        part = RowPartition(n, self.world, self.rank, self.device, self.group)
        self.part = part
        if self.collectives == "native":
            with torch.cuda.device(self.device):
                part.comm = NativeComm(self.group)
        Hd = (FeatureSplitCSRMatrix if self.partition == "feature" else DistCSRMatrix).from_global(Hg, part)
        part.peer = None
        if self.partition == "feature" and self.peer_memory and self.world > 1:
            f_max = max(int(self.hidden_layer_size), out_size)
            fp_max = (-(-f_max // self.world) + 3) // 4 * 4
            try:
                part.peer = PeerBuffers(part, fp_max, rows_max=part.n_loc + part.n_loc // 8 + 64)
            except Exception as e:          # no P2P / IPC on this box: NCCL all-to-all path
                import logging
                logging.getLogger("graphconvgeo_b200").warning("peer-memory transposes unavailable (%s); using NCCL", e)
                part.peer = None
            ok = torch.tensor([1.0 if part.peer is not None else 0.0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if ok.item() < 1.0:
                part.peer = None
        # local rows of X (padded with empty rows)
        ip, ix, d = Xg._host_arrays()
        p = np.zeros(part.n_loc + 1, np.int32)
        p[:part.r1 - part.r0 + 1] = ip[part.r0:part.r1 + 1] - ip[part.r0]
        p[part.r1 - part.r0 + 1:] = p[part.r1 - part.r0]
        self.Xd = CSRMatrix.from_host((p, np.ascontiguousarray(ix[ip[part.r0]:ip[part.r1]]),
                                       np.ascontiguousarray(d[ip[part.r0]:ip[part.r1]])),
                                      (part.n_loc, in_size), self.device, 1024)
        # Decisions about X's layout that shape COLLECTIVES (which terms form the dense head, which rows of X^T are
        # heavy: their gradient blocks are summed over ranks as whole buffers) are taken once from the GLOBAL X, so
        # every rank issues the same all-reduces on buffers of the same shape and meaning.
        import os
        from .sparse import BlockedRows, HeadSplit
        from .lasagne_layers import _is_big
        F = int(self.hidden_layer_size)
        k_head = int(os.environ.get("GCG_X_HEAD", "256"))
        big = _is_big(Xg, F)
        top = None
        if big and k_head > 0 and ops.gemm_uses_tensor_cores(part.n_loc, F, k_head):
            top = HeadSplit.top_terms(Xg.indices, in_size, k_head)
        heavy_ids = None
        if big:
            df = torch.bincount(Xg.indices.to(torch.int64), minlength=in_size)
            if top is not None:
                df[top] = 0                                   # the head terms leave the sparse part
            heavy_ids = BlockedRows.heavy_rows(df.cpu().numpy(), n, F, float(os.environ.get("GCG_XT_BLOCK_MB", "96")),
                                               int(os.environ.get("GCG_XT_HEAVY_FACTOR", "16")))
        self._x_plan = dict(big=big, top=top, heavy_ids=heavy_ids)
        if getattr(self, "keep_host_inputs", False):      # for checkers (host_inputs): the global matrices, model order
            self._host_inputs = (Xg.to_scipy(), Hg.to_scipy())
        del Xg, Hg
        self._build(self.Xd, Hd, in_size, out_size)
        self.l_hid1._x_plan = self._x_plan
        self.ti = {}
        self._sel = {}
        for name, idx in (("train", train_indices), ("dev", dev_indices), ("test", test_indices)):
            t = DistTargetIndices(node_map(idx), Hd, len(idx))
            self.ti[name] = t
            self._sel[name] = t.sel
        self.ti_train = self.ti["train"]
        to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(self.device)
        self.y_train_dev = to_dev(Y[np.asarray(train_indices)][self.ti["train"].sel])
        self.y_dev_dev = to_dev(Y[np.asarray(dev_indices)][self.ti["dev"].sel])
        self._heads = {}
        self._graph = None
        self._steps_done = 0
        self._train_hb = None
        self._red = torch.zeros(3, dtype=torch.float32, device=self.device)
        return self

    # ------------------------------------------------------------------ steps
    def _head(self, logits, y, n_global, grad=None):
        n, C = logits.shape
        hb = self._head_buffers(max(n, 1), C)
        if n > 0:
            ops.softmax_ce(logits, y=y, grad=grad, ce=hb["ce"][:n], hit=hb["hit"][:n], denom=n_global)
            ops.sum_scaled(hb["ce"][:n], 1.0 / n_global, out=hb["out"][0:1])
            ops.sum_scaled(hb["hit"][:n], 1.0 / n_global, out=hb["out"][1:2])
        else:
            hb["out"].zero_()
        return hb

    def _train_step_enqueue(self):
        ti, y = self.ti_train, self.y_train_dev
        logits = self._forward(ti, train=True)
        n, C = logits.shape
        G = self.l_out._mat("G", n, C)
        hb = self._head(logits, y, ti.n_global, grad=G)
        works = [self._all_reduce_async(hb["out"])]
        self._backward(G, works)          # each layer's gradient all-reduce starts as soon as it is computed
        for w in works:
            w.wait()
        self.adam.step()
        self._train_hb = hb

    def _all_reduce_async(self, t):
        """sum over ranks, in place; returns a handle with wait() (a torch Work, or a fence of the native communicator)"""
        comm = self.part.comm
        if comm is not None and t.dtype == torch.float32 and t.is_contiguous():
            return comm.all_reduce(t)
        return dist.all_reduce(t, group=self.group, async_op=True)

    def _backward(self, G, works=None):
        def reduce_grads(ly):
            if works is not None:
                gs = [g for k, g in ly.grads.items()
                      if not (k == "W" and getattr(ly, "_grad_reduce", None) is not None)]   # dW1: summed piece by piece
                comm = self.part.comm                                                        # inside the X^T.dZ1 product
                if comm is not None and gs and all(g.is_contiguous() for g in gs):
                    works.append(comm.all_reduce(gs))          # the layer's gradients in one NCCL group
                else:
                    works.extend(self._all_reduce_async(g) for g in gs)
        # dW1 = X^T.dZ1 is the largest message of the epoch (V x h floats) and the last gradient to be computed:
        # its pieces are all-reduced as they are finished, overlapped with the rest of the product
        self.l_hid1._grad_reduce = self._all_reduce_async if works is not None else None
        from .mlpconv import DropoutLayer
        grad, preact = G, False
        for i in range(len(self.layers) - 1, -1, -1):
            ly = self.layers[i]
            if isinstance(ly, DropoutLayer):          # same rule as the single-GPU loop: the mask scales the
                grad = ly.backward(grad)              # gradient and act' of the layer below is applied there
                preact = False
                continue
            prev = self.layers[i - 1] if i > 0 else None
            mask = None
            if prev is not None and type(prev) in (L.SparseConvolutionDenseLayer, L.ConvolutionDenseLayer) \
                    and prev.nonlinearity in ("rectify", "tanh"):
                mask = (prev._out, prev.nonlinearity)
            if isinstance(ly, L.SparseConvolutionDenseLayer):
                ly.backward(grad, preact=preact)
                reduce_grads(ly)
                break
            if isinstance(ly, L.HighwayConvolutionDenseLayer):
                grad = ly.backward(grad, input_mask=mask)
            else:
                grad = ly.backward(grad, preact=preact, input_mask=mask)
            reduce_grads(ly)
            preact = mask is not None

    def f_train(self):
        self._train_step_enqueue()
        self._steps_done += 1
        return self._train_hb

    def f_val(self, y, ti):
        logits = self._forward(ti, train=False)
        hb = self._head(logits, y, ti.n_global)
        dist.all_reduce(hb["out"], group=self.group)
        reg = self.elastic()
        o = hb["out"].cpu().numpy()
        return float(np.float32(o[0]) + np.float32(reg.item())), float(o[1])

    def _gather_rows_to_all(self, local, ti, width, dtype):
        """assemble per-target rows computed by their owners into the caller's index order"""
        out = torch.zeros((ti.n_global, width), dtype=dtype, device=self.device)
        if ti.n > 0:
            out[torch.from_numpy(ti.sel).to(self.device)] = local.to(dtype).reshape(ti.n, width)
        dist.all_reduce(out, group=self.group)
        return out

    def f_predict(self, ti):
        logits = self._forward(ti, train=False)
        n, C = logits.shape
        hb = self._head_buffers(max(n, 1), C)
        if n > 0:
            ops.softmax_ce(logits, pred=hb["pred"][:n])
        return self._gather_rows_to_all(hb["pred"][:n], ti, 1, torch.int64).cpu().numpy()[:, 0]

    def f_predict_proba(self, ti):
        logits = self._forward(ti, train=False)
        n, C = logits.shape
        probs = self.l_out._mat(("probs", n), max(n, 1), C)
        if n > 0:
            ops.softmax_ce(logits, probs=probs[:n])
        return self._gather_rows_to_all(probs[:n].contiguous(), ti, C, torch.float32).cpu().numpy()

    def accuracy(self, dataset_partition, y_true):
        ti = self._partition(dataset_partition)
        y = torch.from_numpy(np.ascontiguousarray(np.asarray(y_true)[ti.sel], dtype=np.int32)).to(self.device)
        return self.f_val(y, ti)[1]

    def host_inputs(self):
        """the GLOBAL (X, A_hat) in model order; set ``keep_host_inputs = True`` before prepare()"""
        if getattr(self, "_host_inputs", None) is None:
            raise RuntimeError("DistMLPCONV.host_inputs: set keep_host_inputs = True before prepare()")
        return self._host_inputs

    def node_rows(self, t):
        """local slab -> full matrix in ORIGINAL node order (all-gathered; for tests)."""
        part = self.part
        full = torch.zeros((part.n_pad, t.shape[1]), dtype=t.dtype, device=self.device)
        full[part.rank * part.n_loc:(part.rank + 1) * part.n_loc] = t
        dist.all_reduce(full, group=self.group)
        a = full[:part.n_total].cpu().numpy()
        return a if self.node_order is None else a[self.node_inverse]
