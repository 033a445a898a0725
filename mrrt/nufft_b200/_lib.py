"""ctypes binding of libb200nufft.so (C ABI in include/b200nufft.h).

The library is built in-tree by ``python -m mrrt.nufft_b200.build`` (nvcc, sm_100a).
There is no CPU fallback: if the shared object is missing, importing the operator
fails loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200nufft.so")

B2N_OK, B2N_EINVAL, B2N_ECUDA, B2N_ESTATE, B2N_ENONFINITE = 0, 1, 2, 3, 4
B2N_SINGLE, B2N_DOUBLE = 0, 1
B2N_COORD_TM, B2N_COORD_OMEGA = 0, 1

c_int = ctypes.c_int
c_long = ctypes.c_long
c_i64 = ctypes.c_int64
c_vp = ctypes.c_void_p
c_dbl = ctypes.c_double
c_chp = ctypes.c_char_p
PP = ctypes.POINTER(c_vp)
PI = ctypes.POINTER(c_int)

# every symbol include/b200nufft.h declares: (restype, argtypes)
SIGNATURES = {
    "b2n_version": (c_int, []),
    "b2n_last_error": (c_chp, []),
    "b2n_plan_create": (c_int, [c_int, PI, PI, PI, c_int, c_int, c_int, c_int, PP]),
    "b2n_plan_destroy": (c_int, [c_vp]),
    "b2n_plan_set_option": (c_int, [c_vp, c_chp, c_long]),
    "b2n_plan_get_option": (c_long, [c_vp, c_chp]),
    "b2n_plan_set_tables": (c_int, [c_vp, PP]),
    "b2n_plan_set_scaling": (c_int, [c_vp, PP, PP, c_dbl, c_dbl]),
    "b2n_plan_set_points": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp]),
    "b2n_plan_set_sample_phase": (c_int, [c_vp, c_vp, c_vp]),
    "b2n_plan_num_points": (c_i64, [c_vp]),
    "b2n_plan_num_bins": (c_i64, [c_vp]),
    "b2n_plan_get_points": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "b2n_plan_num_slots": (c_i64, [c_vp]),
    "b2n_plan_get_slots": (c_int, [c_vp, c_vp, c_vp]),
    "b2n_interp_fwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_vp]),
    "b2n_interp_adj": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_vp]),
    "b2n_grid_fwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_vp]),
    "b2n_grid_adj": (c_int, [c_vp, c_vp, c_vp, c_int, c_vp]),
    "b2n_grid_multiply": (c_int, [c_vp, c_vp, c_vp, c_int, c_vp]),
    "b2n_nufft_fwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_vp]),
    "b2n_nufft_adj": (c_int, [c_vp, c_vp, c_vp, c_int, c_vp]),
    "b2n_sense_fwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    "b2n_sense_adj": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    "b2n_plan_set_sparse": (c_int, [c_vp, PP, PP, c_i64, c_vp, c_vp]),
    "b2n_plan_sparse_nnz": (c_i64, [c_vp]),
    "b2n_plan_get_sparse": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "b2n_spmv_fwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_vp]),
    "b2n_spmv_adj": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_vp]),
    "b2n_planes_fwd": (c_int, [c_vp, c_vp, c_int, c_int, c_vp, c_vp]),
    "b2n_planes_adj": (c_int, [c_vp, c_vp, c_int, c_int, c_vp, c_vp]),
    "b2n_axis3_fwd": (c_int, [c_vp, c_vp, c_vp]),
    "b2n_axis3_adj": (c_int, [c_vp, c_vp, c_vp]),
    "b2n_slab_scatter": (c_int, [c_vp, c_vp, c_int, c_int, c_int, PP, PI, PI, c_vp]),
    "b2n_slab_gather": (c_int, [c_vp, c_vp, c_int, c_int, c_int, PP, PI, PI, c_vp]),
    "b2n_plan_device_bytes": (c_i64, [c_vp]),
    "b2n_plan_launch_count": (c_i64, [c_vp]),
    "b2n_plan_get_timing": (c_int, [c_vp, ctypes.POINTER(c_dbl)]),
}

_lib = None


def load():
    """Load libb200nufft.so and bind every declared symbol (raises if absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libb200nufft.so is not built (%s). Run `python -m mrrt.nufft_b200.build` "
            "(needs nvcc); there is no CPU fallback." % LIB_PATH
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    """Map a status code to the exception type the reference raises."""
    if rc == B2N_OK:
        return
    msg = load().b2n_last_error().decode("utf-8", "replace")
    if rc in (B2N_EINVAL, B2N_ENONFINITE):
        raise ValueError(msg)
    raise RuntimeError(msg)
