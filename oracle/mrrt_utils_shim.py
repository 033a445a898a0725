"""Minimal stand-in for the un-vendored third-party dependency mrrt.utils
(github.com/mritools/mrrt.utils, version unpinned by the reference)."""
from contextlib import nullcontext
import numpy as np


class _Config:
    have_cupy = False
    have_pyfftw = False


config = _Config()


def profile(f):
    return f


def prod(seq):
    out = 1
    for s in seq:
        out *= int(s)
    return out


def get_array_module(arr, xp=None):
    return np, False


def get_data_address(x):
    return x.__array_interface__["data"][0]


def complexify(x, complex_dtype=None, subok=False):
    x = np.asanyarray(x)
    if complex_dtype is None:
        complex_dtype = np.result_type(x.dtype, np.complex64)
    if x.dtype != complex_dtype:
        x = x.astype(complex_dtype)
    return x


def reale(x, com="error", tol=None, msg=None, xp=None):
    x = np.asanyarray(x)
    if not np.iscomplexobj(x):
        return x
    if tol is None:
        tol = 1000 * np.finfo(x.real.dtype).eps
    mx = np.max(np.abs(x)) if x.size else 0
    if mx == 0:
        return x.real
    frac = np.max(np.abs(x.imag)) / mx
    if frac > tol:
        raise ValueError("imaginary part not negligible: %g" % frac)
    return x.real


def outer_sum(xx, yy):
    xx = np.asanyarray(xx)
    yy = np.asanyarray(yy)
    return xx.reshape(xx.shape + (1,) * yy.ndim) + yy.reshape((1,) * xx.ndim + yy.shape)


def fftn(x, s=None, axes=None, **kw):
    return np.fft.fftn(x, s=s, axes=axes)


def ifftn(x, s=None, axes=None, **kw):
    return np.fft.ifftn(x, s=s, axes=axes)


def max_percent_diff(s1, s2, use_both=False):
    s1 = np.asarray(s1); s2 = np.asarray(s2)
    denom = max(np.abs(s1).max(), np.abs(s2).max()) if use_both else np.abs(s1).max()
    return 100 * np.abs(s1 - s2).max() / denom
