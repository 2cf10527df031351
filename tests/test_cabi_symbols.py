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
