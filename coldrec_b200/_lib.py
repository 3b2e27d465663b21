"""ctypes binding of the C-ABI library (include/coldrec_b200.h).

The library is built in-tree by ``make -C coldrec_b200/csrc`` (``__graft_entry__.build()``).  There is
no CPU fallback: if the shared object is missing, or a compute entry point reports that no sm_100
device is present, the caller gets an exception.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# CR_LIB_PATH: a differently built libcoldrec_b200.so (tools/build_variant.sh, kernel A/B runs); never a fallback
LIB_PATH = os.environ.get("CR_LIB_PATH") or os.path.join(_HERE, "csrc", "libcoldrec_b200.so")

CR_OK = 0
CR_MAX_K = 64
CR_MASK_SCORE = -1.0e9
SCORE_EXACT_F32 = 0
SCORE_TF32_CHECKED = 1
ACT_NONE, ACT_TANH, ACT_LEAKY_RELU = 0, 1, 2

# name -> (restype, argtypes); mirrors include/coldrec_b200.h one to one
_P = c_void_p
SIGNATURES = {
    "cr_strerror": (c_char_p, [c_int]),
    "cr_last_cuda_error": (c_char_p, []),
    "cr_version": (c_int, []),
    "cr_device_check": (c_int, []),
    "cr_launch_count": (ctypes.c_ulonglong, []),
    "cr_profile_enable": (c_int, [c_int]),
    "cr_profile_read": (c_int, [c_int, ctypes.POINTER(c_double), ctypes.POINTER(c_int)]),
    "cr_spmm_plan_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "cr_spmm_plan": (c_int, [_P, c_int64, c_int64, c_int, _P, c_size_t, _P]),
    "cr_spmm_csr_f32": (c_int, [_P, _P, _P, c_int64, c_int64, _P, c_int, _P, _P, _P, c_float, c_float, _P, c_size_t, _P]),
    "cr_spmm_csr_bcast_f32": (c_int, [_P, _P, _P, c_int64, c_int64, _P, c_int, _P, c_int, c_int64, c_int64, c_int64, c_int, _P, _P, _P, _P,
                                      c_float, c_float, _P, c_size_t, _P]),
    "cr_score_topk_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int, c_int]),
    "cr_score_topk_f32": (c_int, [_P, _P, c_int64, _P, _P, c_int64, c_int64, c_int, _P, _P, _P, c_uint8, c_int, _P, _P, _P,
                                  c_int, _P, c_size_t, _P]),
    "cr_debug_tc_tile": (c_int, [_P, c_int64, _P, c_int64, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "cr_debug_tc_timeline": (c_int, [ctypes.POINTER(ctypes.c_ulonglong), c_int]),
    "cr_topk_merge": (c_int, [_P, _P, c_int, c_int64, c_int, _P, _P, _P]),
    "cr_fill_masked": (c_int, [_P, _P, c_int64, c_int, c_int64, _P, c_uint8, _P, _P, _P]),
    "cr_gather_rows_f32": (c_int, [_P, _P, c_int64, c_int, _P, _P]),
    "cr_copy_rows_f32": (c_int, [_P, _P, _P, c_int64, c_int, _P, _P]),
    "cr_rank_metrics_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "cr_rank_metrics": (c_int, [_P, c_int64, c_int, _P, _P, ctypes.POINTER(c_int32), c_int, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "cr_linear_act_f32": (c_int, [_P, c_int64, c_int, _P, c_int64, c_int, _P, c_int64, _P, _P, _P, _P, c_int, c_int, _P,
                                  c_int64, _P, _P]),
    "cr_linear_act_tc_f32": (c_int, [_P, _P, c_int64, c_int, _P, _P, c_int64, c_int, c_int64, _P, _P, c_int64, _P, _P, _P, c_int, c_int, _P,
                                     c_int64, _P, _P, _P, c_int64, _P]),
    "cr_split_tf32": (c_int, [_P, c_int64, c_int64, c_int, _P, _P, c_int64, _P]),
    "cr_topk_rows_f32": (c_int, [_P, c_int64, c_int, c_int64, c_int, _P, c_int, _P, _P, _P]),
    "cr_bn_fold_f32": (c_int, [_P, _P, _P, _P, c_float, c_int, _P, _P, _P]),
    "cr_heater_blend_f32": (c_int, [_P, c_int, _P, _P, c_float, c_float, c_int64, c_int, _P, _P]),
    "cr_bpr_workspace_bytes": (c_size_t, [c_int64]),
    "cr_bpr_fwd_bwd_f32": (c_int, [_P, _P, c_int, _P, _P, _P, c_int64, c_float, _P, _P, _P, _P, c_size_t, _P]),
    "cr_adam_step_f32": (c_int, [_P, _P, _P, _P, c_int64, c_double, c_double, c_double, c_double, c_int64, c_float, _P, _P]),
    "cr_adam_scalars": (c_int, [c_double, c_double, c_double, c_int64, ctypes.POINTER(c_float)]),
    "cr_sample_pairwise": (c_int, [_P, _P, c_int64, _P, _P, c_int32, ctypes.c_uint64, ctypes.c_uint64, c_int64, c_int64, _P, _P, _P,
                                   _P, _P]),
}

_lib = None


class ColdRecB200Error(RuntimeError):
    pass


def load():
    """Load libcoldrec_b200.so (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ColdRecB200Error(
            f"{LIB_PATH} not found: build it with `make -C coldrec_b200/csrc` (or __graft_entry__.build()). "
            "coldrec_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here means header and library disagree
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc, what):
    if rc == CR_OK:
        return
    lib = load()
    msg = lib.cr_strerror(rc).decode()
    if rc == -5:
        msg += ": " + lib.cr_last_cuda_error().decode()
    if rc in (-1, -2, -3, -4):
        raise ValueError(f"{what}: {msg}")
    raise ColdRecB200Error(f"{what}: {msg}")
