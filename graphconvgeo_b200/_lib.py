"""ctypes binding of libgcg.so (the C ABI declared in include/gcg.h).

There is deliberately NO fallback: if the shared object is missing or a call
fails, an exception is raised.  The product path never routes through a CPU or
PyTorch implementation of the hot ops.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgcg.so")

c_i32, c_i64, c_f32, c_vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p
c_int = C.c_int


class GcgError(RuntimeError):
    pass


STATUS = {0: "GCG_OK", -1: "GCG_ERR_BAD_ARG", -2: "GCG_ERR_SHAPE", -3: "GCG_ERR_CUDA",
          -4: "GCG_ERR_NCCL", -5: "GCG_ERR_NOMEM", -6: "GCG_ERR_UNSUPPORTED"}

ACT = {"identity": 0, "linear": 0, None: 0, "rectify": 1, "relu": 1, "tanh": 2, "sigmoid": 3}
GEMM_MODE = {"fma": 0, "tf32x3": 1, "tf32": 2, "tf32x3_chained": 3}

# name -> (restype, argtypes); mirrors include/gcg.h one to one
PROTOTYPES = {
    "gcg_version": (c_int, []),
    "gcg_last_error": (C.c_char_p, []),
    "gcg_launch_count": (c_i64, []),
    "gcg_launch_count_reset": (None, []),
    "gcg_plan_create_csr": (c_int, [c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_i32, C.POINTER(c_vp)]),
    "gcg_plan_destroy": (c_int, [c_vp]),
    "gcg_plan_workspace_bytes": (c_i64, [c_vp, c_i64]),
    "gcg_plan_info": (c_int, [c_vp, C.POINTER(c_i64)]),
    "gcg_spmm_csr_f32": (c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_int, c_int, c_vp, c_i64,
                                 c_vp, c_i64, c_vp, c_i64, c_i32, c_vp, c_i64, c_vp]),
    "gcg_gemm_f32": (c_int, [c_int, c_int, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_f32,
                             c_vp, c_int, c_vp, c_i64, c_int, c_int, c_i32, c_vp, c_i64, c_vp]),
    "gcg_gemm_presplit_f32": (c_int, [c_int, c_int, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_f32,
                                      c_vp, c_int, c_vp, c_i64, c_int, c_int, c_i32, c_vp, c_i64, c_vp,
                                      c_vp, c_vp, c_vp, c_vp]),
    "gcg_tf32_split_f32": (c_int, [c_vp, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "gcg_gemm_workspace_bytes": (c_i64, [c_int, c_int, c_i64, c_i64, c_i64, c_int, c_i32]),
    "gcg_gemm_tc_available": (c_int, []),
    "gcg_colsum_f32": (c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_i64, c_vp]),
    "gcg_colsum_workspace_bytes": (c_i64, [c_i64, c_i64]),
    "gcg_act_bwd_f32": (c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_int, c_vp]),
    "gcg_highway_fwd_f32": (c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp]),
    "gcg_highway_bwd_f32": (c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp,
                                    c_i64, c_vp, c_i64, c_i64, c_i64, c_int, c_vp]),
    "gcg_softmax_ce_f32": (c_int, [c_vp, c_i64, c_vp, c_i64, c_i64, c_f32, c_vp, c_i64, c_vp, c_i64, c_vp,
                                   c_vp, c_vp, c_vp]),
    "gcg_sum_f32": (c_int, [c_vp, c_i64, c_f32, c_vp, c_vp]),
    "gcg_sum_slabs_f32": (c_int, [c_vp, c_i32, c_i64, c_vp, c_vp]),
    "gcg_scatter_rows_f32": (c_int, [c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp]),
    "gcg_gather_rows_f32": (c_int, [c_vp, c_i64, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp]),
    "gcg_put_rows_f32": (c_int, [c_vp, c_i64, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp]),
    "gcg_pack_cols_f32": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i32, c_i64, c_vp, c_vp]),
    "gcg_unpack_cols_f32": (c_int, [c_vp, c_i64, c_i64, c_i32, c_i64, c_vp, c_i64, c_vp]),
    "gcg_adam_step_f32": (c_int, [c_i32, C.POINTER(c_vp), C.POINTER(c_vp), C.POINTER(c_vp), C.POINTER(c_vp),
                                  C.POINTER(c_i64), C.POINTER(c_f32), c_f32, c_f32, c_f32, c_f32, c_vp, c_vp,
                                  c_vp, c_i64, c_vp]),
    "gcg_adam_workspace_bytes": (c_i64, [c_i32, C.POINTER(c_i64)]),
    "gcg_elastic_net_f32": (c_int, [c_i32, C.POINTER(c_vp), C.POINTER(c_i64), C.POINTER(c_f32), c_vp, c_vp,
                                    c_i64, c_vp]),
    "gcg_comm_unique_id": (c_int, [c_vp]),
    "gcg_comm_init": (c_int, [c_vp, c_i32, c_i32, C.POINTER(c_vp)]),
    "gcg_comm_destroy": (c_int, [c_vp]),
    "gcg_comm_info": (c_int, [c_vp, C.POINTER(c_i32), C.POINTER(c_i32)]),
    "gcg_comm_wait": (c_int, [c_vp, c_vp]),
    "gcg_allgather_rows_f32": (c_int, [c_vp, c_vp, c_i64, c_i32, c_vp]),
    "gcg_allreduce_grads_f32": (c_int, [c_vp, c_i32, C.POINTER(c_vp), C.POINTER(c_i64), c_i32, c_vp]),
    "gcg_spmm_rowpart_allgather_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp, c_int,
                                               c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_i32, c_vp, c_i64, c_vp]),
    "gcg_epoch_create": (c_int, [C.POINTER(c_vp)]),
    "gcg_epoch_destroy": (c_int, [c_vp]),
    "gcg_epoch_record_begin": (c_int, [c_vp]),
    "gcg_epoch_record_end": (c_int, [c_vp]),
    "gcg_epoch_size": (c_i64, [c_vp]),
    "gcg_epoch_call_name": (C.c_char_p, [c_vp, c_i64]),
    "gcg_epoch_run": (c_int, [c_vp, c_vp]),
    "gcg_kdtree_fit_host": (c_int, [c_vp, c_i64, c_i32, c_i64, c_vp, C.POINTER(c_i64)]),
    "gcg_ahat_nnz_host": (c_i64, [c_i64, c_vp, c_vp]),
    "gcg_ahat_build_host": (c_int, [c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gcg_ahat_device_workspace_bytes": (c_i64, [c_i64]),
    "gcg_ahat_indptr_device": (c_int, [c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "gcg_ahat_fill_device": (c_int, [c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "gcg_haversine_nearest_f64": (c_int, [c_vp, c_i64, c_vp, c_i32, c_vp, c_vp, c_vp]),
    "gcg_haversine_pairs_f64": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "gcg_csr_transpose_host": (c_int, [c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gcg_csr_gather_rows_host": (c_i64, [c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "gcg_csr_split_colblocks_host": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "gcg_csr_permute_host": (c_int, [c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gcg_peer_alloc": (c_int, [c_i64, C.POINTER(c_vp), c_vp]),
    "gcg_peer_free": (c_int, [c_vp]),
    "gcg_peer_open": (c_int, [c_vp, C.POINTER(c_vp)]),
    "gcg_peer_close": (c_int, [c_vp]),
    "gcg_push_cols_f32": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i32, c_i64, C.POINTER(c_vp), c_i64, c_vp]),
    "gcg_push_rows_f32": (c_int, [c_vp, C.POINTER(c_i64), c_i32, c_i64, C.POINTER(c_vp), c_i64, c_vp]),
    "gcg_peer_barrier": (c_int, [C.POINTER(c_vp), c_vp, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "gcg_spmm_csr_routed_f32": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i32, C.POINTER(c_vp), C.POINTER(c_i64), c_i64, c_vp,
                                        c_int, c_i32, c_vp, c_i64, c_vp]),
    "gcg_spmm_set_tuning": (None, [c_int, c_int]),
    "gcg_spmm_set_group": (None, [c_int, c_int]),
    "gcg_plan_set_schedule": (c_int, [c_vp, c_i64, c_vp, c_vp]),
    "gcg_spmm_stream_tuning": (None, [c_int, c_int, c_int]),
    "gcg_plan_set_near_window": (c_int, [c_vp, c_i32]),
    "gcg_spgemm_workspace_bytes": (c_i64, [c_i64]),
    "gcg_spgemm_count_csr": (c_int, [c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_i64, c_vp]),
    "gcg_spgemm_fill_pattern_csr": (c_int, [c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "gcg_spgemm_fill_csr_f32": (c_int, [c_i64, c_i64, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                        c_vp, c_i64, c_vp]),
    "gcg_csr_gather_rows_device": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "gcg_csr_transpose_device_workspace_bytes": (c_i64, [c_i64, c_i64, c_i64]),
    "gcg_csr_transpose_device": (c_int, [c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
}

_lib = None


def lib():
    """The loaded libgcg.so; raises (never falls back) when it is unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "graphconvgeo_b200: %s is missing. Build it with `python -m graphconvgeo_b200.build` "
            "(needs nvcc, targets sm_100a). There is no CPU/PyTorch fallback for the hot path." % LIB_PATH)
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        try:
            fn = getattr(handle, name)
        except AttributeError:
            continue            # optional symbol groups are checked by tests/test_abi.py
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().gcg_last_error()
        raise GcgError("%s failed: %s (%s)" % (what or "libgcg call", STATUS.get(status, status),
                                               msg.decode() if msg else ""))


def act_code(name) -> int:
    if isinstance(name, int):
        return name
    if name not in ACT:
        raise ValueError("unsupported nonlinearity %r" % (name,))
    return ACT[name]
