"""ctypes binding of libgpet_b200.so (the C ABI declared in include/gpet_b200.h).

There is no CPU fallback: if the shared library is missing every entry point raises. The library
is built in-tree by `gaussian_process_edge_trace_b200/csrc/build.sh` (see `__graft_entry__.build`).
"""
import ctypes
import os
from ctypes import c_double, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpet_b200.so")

_P = c_void_p  # every device pointer travels as a plain address
ABI_VERSION = 8  # GPET_ABI_VERSION of include/gpet_b200.h this binding was written against

# name -> (restype, argtypes); mirrors include/gpet_b200.h one to one
SIGNATURES = {
    "gpet_last_error": (ctypes.c_char_p, []),
    "gpet_abi_version": (c_int, []),
    "gpet_set_tuning": (c_int, [c_int, c_int]),
    "gpet_comp_grad_img_f64": (c_int, [_P, c_int, c_int, c_int, _P, c_int, c_int, _P, _P, _P]),
    "gpet_comp_grad_img_fast_f32": (c_int, [_P, c_int, c_int, c_int, _P, c_int, c_int, _P, _P, _P]),
    "gpet_normalise_f32": (c_int, [_P, c_int, c_int, c_int, _P, _P]),
    "gpet_grad_kde_workspace_bytes": (c_int64, [c_int, c_int, c_int]),
    "gpet_grad_kde_f32": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P]),
    "gpet_transpose_f32": (c_int, [_P, c_int, c_int, c_int, _P, _P]),
    "gpet_posterior_lowrank_workspace_bytes": (c_int64, [c_int, c_int, c_int]),
    "gpet_posterior_lowrank_f64": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, c_double, c_double, _P, _P, _P,
                                           c_int, _P, _P, _P, _P, _P, _P]),
    "gpet_posterior_full_workspace_bytes": (c_int64, [c_int, c_int, c_int]),
    "gpet_posterior_full_f64": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P, c_double, c_double, _P, _P, _P, _P,
                                        _P, _P, _P]),
    "gpet_sym_eig_workspace_bytes": (c_int64, [c_int, c_int]),
    "gpet_sym_eig_f64": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P]),
    "gpet_factor_assemble_f64": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P, _P]),
    "gpet_sample_f64": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "gpet_sample_score_supported": (c_int, [c_int, c_int, c_int]),
    "gpet_sample_score_f64": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "gpet_sample_keep_f64": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "gpet_standard_normal_workspace_bytes": (c_int64, [c_int64, c_int]),
    "gpet_standard_normal_fixup_bytes": (c_int64, [c_int64, c_int]),
    "gpet_standard_normal_t_f64": (c_int, [ctypes.c_uint32, c_int64, c_int, c_int, c_int64, c_int64, _P, _P, _P, _P, _P]),
    "gpet_host_log_f64": (c_int, [_P, _P, c_int64]),
    "gpet_standard_normal_fixup_apply_f64": (c_int, [c_int64, c_int, c_int, c_int64, c_int64, _P, _P, c_int64, _P]),
    "gpet_score_f64": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "gpet_topk_f64": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, _P]),
    "gpet_density_workspace_bytes": (c_int64, [c_int, c_int, c_int, c_int]),
    "gpet_density_f64": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P]),
    "gpet_density_splat_f64": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "gpet_density_finish_f64": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P]),
    "gpet_select_f64": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P, _P, c_int, _P, _P, c_int, c_int, _P, _P, _P]),
    "gpet_density_bands_supported": (c_int, [c_int, c_int, c_int]),
    "gpet_density_bands_workspace_bytes": (c_int64, [c_int, c_int, c_int]),
    "gpet_density_bands_f64": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P, _P,
                                       _P, _P, _P]),
    "gpet_select_bands_f64": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P, c_int, _P, _P, c_int, c_int, _P, _P,
                                      _P]),
    "gpet_kde_bands_f32": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "gpet_kde_normalised_f32": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P]),
    "gpet_lml_f64": (c_int, [_P, _P, _P, _P, _P, c_int, _P, _P, c_int, c_int, c_double, _P, _P, _P]),
    "gpet_lbfgsb_state_doubles": (c_int64, []),
    "gpet_lbfgsb_state_ints": (c_int64, []),
    "gpet_lbfgsb_init_f64": (c_int, [_P, _P, c_int, _P, _P, _P, _P]),
    "gpet_lbfgsb_advance_f64": (c_int, [_P, _P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P]),
    "gpet_fit_rounds_f64": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_double, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P,
                                    _P, _P, _P, _P]),
    "gpet_lbfgsb_result_f64": (c_int, [_P, _P, c_int, _P, _P, _P, _P, _P]),
    "gpet_lbfgsb_host_init": (c_int, [_P, _P, c_int, _P, _P, _P]),
    "gpet_lbfgsb_host_advance": (c_int, [_P, _P, c_int, _P, _P, _P, _P, _P]),
    "gpet_update_obs_f64": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P]),
    "gpet_training_sets_f64": (c_int, [_P, _P, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P,
                                       _P, _P, _P, _P, _P, _P]),
    "gpet_test_img_f64": (c_int, [_P, _P, c_int, c_int, c_int, c_double, c_int, _P, c_double, _P, _P]),
    "gpet_trace_metrics_f64": (c_int, [_P, _P, c_int, c_int, _P, _P]),
    "gpet_final_predict_f64": (c_int, [_P, _P, _P, _P, c_int, c_int, _P, c_int, c_double, _P, c_int, _P, _P, _P, _P,
                                       _P]),
    "gpet_lml_big_workspace_bytes": (c_int64, [c_int, c_int]),
    "gpet_lml_big_f64": (c_int, [_P, _P, _P, _P, c_int, _P, _P, c_int, c_int, c_double, _P, _P, _P, c_int64, _P]),
    "gpet_fit_rounds_big_f64": (c_int, [_P, _P, _P, _P, c_int, c_int, c_double, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P,
                                        _P, _P, _P, _P, c_int64, _P]),
    "gpet_final_predict_big_workspace_bytes": (c_int64, [c_int, c_int, c_int]),
    "gpet_final_predict_big_f64": (c_int, [_P, _P, _P, _P, c_int, c_int, _P, c_int, c_double, _P, c_int, _P, _P, _P, _P,
                                           _P, _P]),
    "gpet_block_jacobi_workspace_bytes": (c_int64, [c_int, c_int]),
    "gpet_block_jacobi_init_f64": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P]),
    "gpet_block_jacobi_warm_f64": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, _P]),
    "gpet_block_jacobi_sweep_f64": (c_int, [_P, _P, c_int, c_int, _P, _P, _P]),
    "gpet_block_jacobi_factor_f64": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P]),
    "gpet_dense_potrf_f64": (c_int, [_P, c_int, _P, c_int, c_int, _P, _P]),
    "gpet_dense_trsm_f64": (c_int, [_P, c_int, _P, c_int, c_int, _P, c_int, c_int, _P]),
}


class GpetError(RuntimeError):
    pass


_lib = None


def load():
    """Loads the shared library (once). Raises GpetError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GpetError(
            f"{LIB_PATH} is missing: build it with gaussian_process_edge_trace_b200/csrc/build.sh "
            "(there is no CPU fallback for the B200 hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    built = int(lib.gpet_abi_version())
    if built != ABI_VERSION:
        raise GpetError(f"{LIB_PATH} was built from ABI version {built}, this binding needs {ABI_VERSION}: rebuild it with "
                        "gaussian_process_edge_trace_b200/csrc/build.sh")
    # GPET_TUNE="knob=value,knob=value": launch-shape / variant knobs of gpet_set_tuning for experiments (defaults are the
    # measured best; see GPET_TUNE_* in include/gpet_b200.h)
    for item in filter(None, os.environ.get("GPET_TUNE", "").split(",")):
        knob, _, value = item.partition("=")
        if lib.gpet_set_tuning(int(knob), int(value)) != 0:
            raise GpetError(f"GPET_TUNE={item!r}: " + lib.gpet_last_error().decode("utf-8", "replace"))
    _lib = lib
    return lib


def ptr(t):
    """Device (or host) address of a torch tensor / None."""
    if t is None:
        return None
    return t.data_ptr()


def call(name, *args):
    """Calls an int-returning entry point and raises GpetError on a non-zero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.gpet_last_error().decode("utf-8", "replace")
        raise GpetError(f"{name} failed (code {rc}): {msg}")
    return rc


def query(name, *args):
    return getattr(load(), name)(*args)
