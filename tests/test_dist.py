"""Multi-process path: world_size-2 gloo test of the host logic on CPU, and (on a box with >= 2
GPUs) the NCCL row-partitioned model against the oracle."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_row_partition_host_logic_gloo_world2(built_lib):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import dist_worker
    mp.spawn(dist_worker.cpu_main, args=(2, free_port()), nprocs=2, join=True)


def test_row_partition_uneven_and_padding():
    from graphconvgeo_b200.dist import RowPartition, split_columns
    import numpy as np
    for n, world in ((10, 4), (1000, 8), (7, 2), (1400000, 8)):
        parts = [RowPartition(n, world, r) for r in range(world)]
        assert sum(p.r1 - p.r0 for p in parts) == n
        assert all(p.n_loc == parts[0].n_loc and p.n_loc % 4 == 0 for p in parts)
        assert parts[-1].r1 == n and parts[0].r0 == 0
        rows = np.arange(n)
        owners = parts[0].owner(rows)
        for p in parts:
            sel, loc = p.local_rows(rows)
            assert np.array_equal(sel, np.flatnonzero(owners == p.rank))
            assert np.array_equal(loc, rows[sel] - p.rank * p.n_loc)
    ip = np.array([0, 2, 2, 5], np.int32)
    ix = np.array([0, 7, 1, 4, 9], np.int32)
    d = np.arange(5, dtype=np.float32)
    (dp, di, dd), (op, oi, od) = split_columns((ip, ix, d), 4, 8)
    assert list(dp) == [0, 1, 1, 2] and list(di) == [7, 4] and list(dd) == [1, 3]
    assert list(op) == [0, 1, 1, 3] and list(oi) == [0, 1, 9]


@pytest.mark.gpu
def test_row_partitioned_model_matches_oracle_nccl_world2():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_GPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
