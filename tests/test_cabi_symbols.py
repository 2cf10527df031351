"""CPU test: libgpet_b200.so loads without a GPU and exports every symbol include/gpet_b200.h declares,
and the ctypes table in _cabi.py covers exactly those symbols."""
import ctypes
import os
import re

from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "gpet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gpet_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from gaussian_process_edge_trace_b200 import _cabi
    assert os.path.exists(_cabi.LIB_PATH)
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in gpet_b200.h but not exported"
    assert sorted(_cabi.SIGNATURES) == syms
    loaded = _cabi.load()
    assert loaded.gpet_abi_version() == _cabi.ABI_VERSION == 8
    # argument validation happens before any CUDA call, so it is testable without a GPU
    rc = loaded.gpet_score_f64(None, None, None, 1, 500, 1000, 500, 500, 0, None, None)
    assert rc == 1 and b"gpet_score_f64" in loaded.gpet_last_error()
    assert loaded.gpet_density_workspace_bytes(2, 10, 10, 5) >= 2 * 10 * 10 * 8


def test_workspace_queries_and_argument_checks_without_a_gpu():
    """The size queries of the HBM-resident paths are plain host arithmetic and the entry points validate their arguments
    before any CUDA call, so both are testable here."""
    import __graft_entry__
    __graft_entry__.build()
    from gaussian_process_edge_trace_b200 import _cabi
    lib = _cabi.load()
    # posterior: shared-memory kernels up to GPET_MAX_TRAIN = 224, beyond that the training matrices live in the workspace
    small = lib.gpet_posterior_full_workspace_bytes(4, 224, 500)
    big = lib.gpet_posterior_full_workspace_bytes(4, 225, 500)
    assert small == 4 * 224 * 500 * 8 + 4 * 2 * 8 + 256
    assert big >= 4 * (256 * 256 + 256 * 512) * 8                      # K (ld = 256) + K*^T (ld x 512) per trace
    assert lib.gpet_posterior_lowrank_workspace_bytes(4, 2046, 76) >= 4 * (2048 * 2048 + 2048 * 128) * 8
    one = lib.gpet_lml_big_workspace_bytes(1, 2046)
    assert one >= 2 * 2048 * 2048 * 8 and lib.gpet_lml_big_workspace_bytes(208, 2046) >= 208 * (one - 2048)
    assert lib.gpet_final_predict_big_workspace_bytes(16, 2046, 4096) >= 16 * (2048 * 2048 + 2048 * 4096) * 8
    assert lib.gpet_block_jacobi_workspace_bytes(16, 4096) > 0 and lib.gpet_block_jacobi_workspace_bytes(16, 4000) == 0
    rc = lib.gpet_block_jacobi_sweep_f64(None, None, 1, 128, None, None, None)
    assert rc == 1 and b"gpet_block_jacobi_sweep_f64" in lib.gpet_last_error()
    rc = lib.gpet_dense_potrf_f64(None, 100, None, 1, 64, None, None)       # ld must be a multiple of 64
    assert rc == 1 and b"gpet_dense_potrf_f64" in lib.gpet_last_error()
    rc = lib.gpet_lml_big_f64(None, None, None, None, 300, None, None, 1, 0, 1e-6, None, None, None, 0, None)
    assert rc == 1 and b"gpet_lml_big_f64" in lib.gpet_last_error()
    assert lib.gpet_set_tuning(13, 1) == 0 and lib.gpet_set_tuning(99, 1) == 1
