"""Peer-write bandwidth probe for the feature-sliced transposes.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 scripts/p2p_probe.py

Every rank pushes a fixed number of bytes into its peers' IPC-mapped buffers with gcg_push_rows_f32 for a
sweep of run lengths (Fp floats per row: the transposes write one Fp*4-byte run per row), to one peer and
spread over all peers.  Timing: CUDA events on the launching stream, max over ranks.  The question it
answers: do 300-600-byte runs cap the NVLink write rate (then pack + push long runs), or is the cap elsewhere?
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from graphconvgeo_b200 import _lib  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    total_floats = 64 << 20                                   # 256 MB pushed per measurement
    own = C.c_void_p()
    h = C.create_string_buffer(64)
    _lib.check(L.gcg_peer_alloc(total_floats * 4, C.byref(own), h), "gcg_peer_alloc")
    handles = [None] * world
    dist.all_gather_object(handles, h.raw)
    ptrs = []
    for q in range(world):
        if q == rank:
            ptrs.append(own.value)
            continue
        p = C.c_void_p()
        _lib.check(L.gcg_peer_open(C.create_string_buffer(handles[q], 64), C.byref(p)), "gcg_peer_open")
        ptrs.append(p.value)
    src = torch.randn(total_floats, dtype=torch.float32, device=dev)
    flag = torch.zeros(1, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    peer_arr = (C.c_void_p * world)(*ptrs)
    results = []
    for fp in (76, 152, 300, 600, 2400, 16384, 1 << 20):
        n_rows = total_floats // fp
        for mode in ("one_peer", "all_peers"):
            off = [0] * (world + 1)
            if mode == "one_peer":
                tgt = (rank + 1) % world
                for q in range(world + 1):
                    off[q] = 0 if q <= tgt else n_rows
            else:                                             # own share stays local, like the real transposes
                for q in range(world + 1):
                    off[q] = n_rows * q // world
            row_off = (C.c_int64 * (world + 1))(*off)
            times = []
            for it in range(6):
                dist.all_reduce(flag)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(L.gcg_push_rows_f32(src.data_ptr(), row_off, world, fp, peer_arr, 0, stream), "push")
                e1.record()
                torch.cuda.synchronize()
                if it:
                    times.append(e0.elapsed_time(e1))
            t = torch.tensor([min(times)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            remote = n_rows * fp * 4 * (1.0 if mode == "one_peer" else (world - 1) / world)
            results.append(dict(run_bytes=fp * 4, mode=mode, ms=float(t.item()),
                                remote_GBps=remote / (float(t.item()) * 1e-3) / 1e9))
    # the forward transpose of the epoch: gcg_push_cols_f32 on a [n_loc, 600] operand, every rank to every peer
    F = 600
    fp = (-(-F // world) + 3) // 4 * 4
    n_loc = total_floats // (world * fp)
    z = torch.randn(n_loc, F, dtype=torch.float32, device=dev)
    times = []
    for it in range(6):
        dist.all_reduce(flag)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.gcg_push_cols_f32(z.data_ptr(), F, n_loc, F, world, fp, peer_arr, rank * n_loc, stream), "push_cols")
        e1.record()
        torch.cuda.synchronize()
        if it:
            times.append(e0.elapsed_time(e1))
    t = torch.tensor([min(times)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    remote = n_loc * fp * 4 * (world - 1)
    results.append(dict(kernel="push_cols", n_loc=n_loc, F=F, Fp=fp, ms=float(t.item()),
                        remote_GBps=remote / (float(t.item()) * 1e-3) / 1e9))
    if rank == 0:
        print(json.dumps(dict(world=world, bytes_per_push=total_floats * 4, results=results)))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
