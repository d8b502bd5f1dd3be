"""The C-ABI library loads and exports every symbol include/gcg.h declares (no GPU needed)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "gcg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gcg_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for must in ["gcg_spmm_csr_f32", "gcg_gemm_f32", "gcg_softmax_ce_f32", "gcg_adam_step_f32",
                 "gcg_highway_bwd_f32", "gcg_kdtree_fit_host", "gcg_plan_create_csr", "gcg_last_error",
                 # widened rows: A_hat on the device, smoothing SpGEMM, minibatch slicing,
                 # graph projection, label pipeline
                 "gcg_ahat_fill_device", "gcg_spgemm_count_csr", "gcg_spgemm_fill_csr_f32",
                 "gcg_spgemm_fill_pattern_csr", "gcg_csr_gather_rows_device", "gcg_csr_transpose_device",
                 "gcg_haversine_nearest_f64"]:
        assert must in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, "declared in gcg.h but not exported: %s" % missing


def test_python_prototypes_cover_the_header(built_lib):
    from graphconvgeo_b200 import _lib
    decl = set(declared_symbols())
    proto = set(_lib.PROTOTYPES)
    assert decl <= proto, "no ctypes prototype for: %s" % sorted(decl - proto)
    assert _lib.lib().gcg_version() >= 100


def test_error_convention(built_lib):
    """bad arguments return a negative status and leave a message (no CUDA call involved)."""
    from graphconvgeo_b200 import _lib
    L = _lib.lib()
    rc = L.gcg_spmm_csr_f32(None, None, 0, 0, None, 0, None, 0, 0, None, 0, None, 0, None, 0, 0, None, 0, None)
    assert rc == -1
    assert b"plan is NULL" in L.gcg_last_error()
    try:
        _lib.check(rc, "gcg_spmm_csr_f32")
    except _lib.GcgError as e:
        assert "GCG_ERR_BAD_ARG" in str(e)
    else:
        raise AssertionError("check() must raise")
    # argument checks of the widened rows fire before any CUDA call as well
    assert L.gcg_spgemm_count_csr(4, 4, None, None, None, None, 0, None, None, 0, None) == -1
    assert L.gcg_spgemm_fill_csr_f32(4, 4, None, None, None, 1, None, None, None, None, None, None, None, 0, None) == -1
    assert b"c_indptr is NULL" in L.gcg_last_error()
    assert L.gcg_csr_transpose_device(-1, 4, 0, None, None, None, None, None, None, None, 0, None) == -1
    assert L.gcg_spgemm_workspace_bytes(9000) > 9000 * 8 * 148


def test_no_cpu_fallback_in_product_path():
    """ops refuse CPU tensors loudly; the package never imports the oracle."""
    import torch
    from graphconvgeo_b200 import ops
    import pytest
    with pytest.raises(TypeError):
        ops.colsum(torch.zeros(4, 4))
    for fn in os.listdir(os.path.join(ROOT, "graphconvgeo_b200")):
        if fn.endswith(".py"):
            txt = open(os.path.join(ROOT, "graphconvgeo_b200", fn)).read()
            assert "import oracle" not in txt and "from oracle" not in txt, fn


def test_epoch_program_object_without_a_gpu():
    """gcg_epoch_*  bookkeeping needs no device: create / record window / run of an
    empty program / error statuses and messages."""
    import ctypes as C
    from graphconvgeo_b200 import _lib
    L = _lib.lib()
    h, h2 = C.c_void_p(), C.c_void_p()
    assert L.gcg_epoch_create(C.byref(h)) == 0 and L.gcg_epoch_create(C.byref(h2)) == 0
    assert L.gcg_epoch_size(h) == 0 and L.gcg_epoch_run(h, None) == 0
    assert L.gcg_epoch_record_begin(h) == 0
    assert L.gcg_epoch_run(h, None) == -1 and b"still being recorded" in L.gcg_last_error()
    assert L.gcg_epoch_record_begin(h2) == -1 and b"already recording" in L.gcg_last_error()
    assert L.gcg_epoch_record_end(h2) == -1
    assert L.gcg_epoch_record_end(h) == 0 and L.gcg_epoch_run(h, None) == 0
    assert L.gcg_epoch_call_name(h, 0) == b"" and L.gcg_epoch_call_name(None, 0) == b""
    assert L.gcg_epoch_create(None) == -1
    assert L.gcg_epoch_destroy(h) == 0 and L.gcg_epoch_destroy(h2) == 0 and L.gcg_epoch_destroy(None) == 0


def test_multi_gpu_entry_points_check_their_arguments():
    """gcg_comm_* / gcg_allreduce_grads_f32 / gcg_spmm_rowpart_allgather_f32: argument errors are
    reported before NCCL or CUDA are touched (NCCL is only bound, by dlopen, inside gcg_comm_unique_id / _init)."""
    import ctypes as C
    from graphconvgeo_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    assert L.gcg_comm_init(None, 2, 0, C.byref(h)) == -1 and b"NULL" in L.gcg_last_error()
    uid = C.create_string_buffer(128)
    assert L.gcg_comm_init(uid, 2, 2, C.byref(h)) == -1 and b"rank 2 of 2" in L.gcg_last_error()
    assert L.gcg_comm_unique_id(None) == -1
    assert L.gcg_comm_wait(None, None) == -1
    assert L.gcg_allreduce_grads_f32(None, 0, None, None, 1, None) == -1
    assert L.gcg_allgather_rows_f32(None, None, 0, 1, None) == -1
    assert L.gcg_spmm_rowpart_allgather_f32(None, None, None, None, 0, 0, 0, None, 0, None, 0, None, 0, None, 0, None, 0,
                                            0, None, 0, None) == -1
    assert L.gcg_comm_destroy(None) == 0
