"""The ctypes binding a maintainer of mrrt.nufft would add to the reference as
``mrrt/nufft/_b200.py``: it replaces ``NufftBase._init_gpu``
(_nufft.py:362-390) and the GPU branches of ``_nufft_table_interp`` (:1057-1084) and
``_nufft_table_adj`` (:1166-1189) with calls into ``libb200nufft.so``.  Coordinates are handed
over as the reference's own ``obj.tm`` (``B2N_COORD_TM``) and tables as its own ``obj.h``.

In the reference the device arrays are CuPy arrays (``CupyArrays``).  CuPy is not installed in
this repository's image, so ``tests/test_abi_stub.py`` executes the very same three functions
with ``TorchArrays`` (PyTorch tensors as the device-array type); nothing else differs.
"""
import ctypes
import os

import numpy as np

_vp, _i, _i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
_lib = None


def load(path=None):
    global _lib
    if _lib is None:
        path = path or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    "mrrt", "nufft_b200", "libb200nufft.so")
        lib = ctypes.CDLL(path)
        lib.b2n_plan_create.argtypes = [_i, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i),
                                        _i, _i, _i, _i, ctypes.POINTER(_vp)]
        lib.b2n_plan_destroy.argtypes = [_vp]
        lib.b2n_plan_set_tables.argtypes = [_vp, ctypes.POINTER(_vp)]
        lib.b2n_plan_set_points.argtypes = [_vp, _vp, _i64, _i, _vp]
        lib.b2n_interp_fwd.argtypes = [_vp, _vp, _vp, _i, _i, _vp]
        lib.b2n_interp_adj.argtypes = [_vp, _vp, _vp, _i, _i, _vp]
        lib.b2n_last_error.restype = ctypes.c_char_p
        _lib = lib
    return _lib


def _check(rc):
    if rc:
        msg = _lib.b2n_last_error().decode()
        raise (ValueError if rc in (1, 4) else RuntimeError)(msg)


class CupyArrays(object):
    """Device arrays as the reference has them."""

    def __init__(self):
        import cupy

        self.cupy = cupy

    def device_id(self):
        return self.cupy.cuda.Device().id

    def stream_ptr(self):
        return self.cupy.cuda.get_current_stream().ptr

    def to_host(self, a):
        return self.cupy.asnumpy(a)

    def fortran(self, a):
        return self.cupy.asfortranarray(self.cupy.asarray(a))

    def empty_f(self, shape, dtype, zero=False):
        return (self.cupy.zeros if zero else self.cupy.empty)(shape, dtype=dtype, order="F")

    def ptr(self, a):
        return a.data.ptr


class TorchArrays(object):
    """The same operations on PyTorch CUDA tensors (2-D Fortran order = transposed view of a
    C-contiguous tensor)."""

    def __init__(self):
        import torch

        self.torch = torch

    def device_id(self):
        return self.torch.cuda.current_device()

    def stream_ptr(self):
        return self.torch.cuda.current_stream().cuda_stream

    def to_host(self, a):
        return a.cpu().numpy() if isinstance(a, self.torch.Tensor) else np.asarray(a)

    def fortran(self, a):
        t = a if isinstance(a, self.torch.Tensor) else self.torch.from_numpy(np.asarray(a))
        t = t.cuda()
        if t.dim() == 1:
            return t.contiguous()
        return t.t().contiguous().t()

    def empty_f(self, shape, dtype, zero=False):
        tdt = {np.dtype(np.complex64): self.torch.complex64,
               np.dtype(np.complex128): self.torch.complex128}[np.dtype(dtype)]
        mk = self.torch.zeros if zero else self.torch.empty
        return mk(tuple(reversed(shape)), dtype=tdt, device="cuda").t()

    def ptr(self, a):
        return a.data_ptr()


def init_gpu(obj, xp):                  # replaces NufftBase._init_gpu (_nufft.py:362-390)
    load()
    a3 = lambda v: (_i * 3)(*(list(v) + [1] * (3 - len(v))))
    plan = _vp()
    _check(_lib.b2n_plan_create(obj.ndim, a3(obj.Nd), a3(obj.Kd), a3(obj.Jd), obj.Ld,
                                0 if obj.precision == "single" else 1,
                                int(obj.phasing == "complex"),
                                xp.device_id(), ctypes.byref(plan)))
    hs = [np.ascontiguousarray(xp.to_host(h)) for h in obj.h]     # tables: host pointers
    ptrs = (_vp * 3)(*([h.ctypes.data for h in hs] + [None] * (3 - len(hs))))
    _check(_lib.b2n_plan_set_tables(plan, ptrs))
    tm = xp.fortran(obj.tm)                                       # [M, ndim] column-major
    stream = _vp(xp.stream_ptr())
    _check(_lib.b2n_plan_set_points(plan, xp.ptr(tm), obj.M, 0, stream))   # 0 = B2N_COORD_TM
    obj._b200_plan = plan


def table_interp(obj, xk, xp):          # replaces the GPU branch at _nufft.py:1057-1084
    reps = xk.shape[-1]
    xk = xp.fortran(xk)                 # [prod(Kd), reps], first axis fastest
    x = xp.empty_f((obj.M, reps), obj._cplx_dtype, zero=True)
    stream = _vp(xp.stream_ptr())
    _check(_lib.b2n_interp_fwd(obj._b200_plan, xp.ptr(xk), xp.ptr(x), reps, 0, stream))
    return x


def table_adj(obj, x, xp):              # replaces the GPU branch at _nufft.py:1166-1189
    reps = x.shape[-1]
    x = xp.fortran(x)
    xk = xp.empty_f((int(np.prod(obj.Kd)), reps), obj._cplx_dtype)
    stream = _vp(xp.stream_ptr())
    _check(_lib.b2n_interp_adj(obj._b200_plan, xp.ptr(x), xp.ptr(xk), reps, 0, stream))
    return xk                           # the library zeroes xk first, like the C code


def destroy(obj):
    plan = getattr(obj, "_b200_plan", None)
    if plan is not None:
        _lib.b2n_plan_destroy(plan)
        obj._b200_plan = None
