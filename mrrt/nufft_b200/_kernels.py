"""Interpolation-kernel mathematics evaluated on the HOST at plan time.

These are small 1-D computations (a few thousand values per axis); they stay in NumPy /
SciPy and only their results are uploaded.  Public names mirror the reference:
``BeattyKernel``, ``KaiserBesselKernel``, ``kaiser_bessel``, ``kaiser_bessel_ft``
(mrrt/nufft/_kernels.py, mrrt/nufft/_kaiser_bessel.py).
"""
from math import sqrt

import numpy as np
from scipy.special import i0, iv, jv

__all__ = ["BeattyKernel", "KaiserBesselKernel", "kaiser_bessel", "kaiser_bessel_ft"]


def kaiser_bessel(x, J=6, alpha=None, m=0):
    """Generalised Kaiser-Bessel window on the support ``[-J/2, J/2]``.

    ``KB(x) = f^m I_m(alpha f) / I_m(alpha)``, ``f = sqrt(1 - (2x/J)^2)``; zero outside
    the support (reference: _kaiser_bessel.py:78-149).
    """
    x = np.asarray(x, dtype=np.float64)
    if alpha is None:
        alpha = 2.34 * J
    order = abs(m)
    out = np.zeros(x.shape, dtype=np.float64)
    inside = 2 * np.abs(x) < J
    r = 2 * x[inside] / J
    f = np.sqrt(1 - r * r)
    if order == 0:
        out[inside] = i0(alpha * f) / float(i0(alpha))
    else:
        out[inside] = f ** m * iv(order, alpha * f) / float(iv(order, alpha))
    return out


def kaiser_bessel_ft(u, J=6, alpha=None, m=0, d=1):
    """Fourier transform of :func:`kaiser_bessel` (Lewitt 1990, eq. A3).

    Reference: _kaiser_bessel.py:153-226.  A complex square root keeps the formula
    valid past the main lobe; the real part is returned.
    """
    u = np.asarray(u, dtype=np.float64)
    if not alpha:
        alpha = 2.34 * J
    q = (np.pi * J) * u
    q = q * q - alpha * alpha
    z = np.lib.scimath.sqrt(q)   # complex where the argument is negative
    nu = d / 2.0 + m
    amp = (2 * np.pi) ** (d / 2.0) * (J / 2.0) ** d * alpha ** m
    amp = amp / (i0(alpha) if m == 0 else iv(m, alpha))
    y = amp * jv(nu, z) / z ** nu
    return np.real(y)


class KaiserBesselKernel(object):
    """Separable KB kernel with user-supplied shape parameters ("kb:user")."""

    kernel_type = "kb:user"

    def __init__(self, shape, alpha=None, m=None):
        if np.isscalar(shape):
            shape = (shape,)
        self.shape = tuple(int(s) for s in shape)
        if alpha is None or m is None:
            raise ValueError("kwargs must contain shape, m, alpha for kb:user case")
        self.alpha = list(np.atleast_1d(alpha).astype(float))
        self.m = list(np.atleast_1d(m).astype(float))
        if len(self.alpha) != self.ndim or len(self.m) != self.ndim:
            raise ValueError("array did not have the expected size of {}".format(self.ndim))
        self.params = {}
        self._make_kernels()

    @property
    def ndim(self):
        return len(self.shape)

    def _make_kernels(self):
        self.is_kaiser_scale = True
        self.kernels = []
        for J, a, m in zip(self.shape, self.alpha, self.m):
            self.kernels.append(lambda x, J=J, a=a, m=m: kaiser_bessel(x, J=J, alpha=a, m=m))

    def __str__(self):
        return "kernel type: {}\nkernel shape: {}\n".format(self.kernel_type, self.shape)


class BeattyKernel(KaiserBesselKernel):
    """KB kernel with the shape parameter of Beatty et al., IEEE TMI 24(6), eq. 5:
    ``alpha = pi sqrt(J^2/(K/N)^2 (K/N - 1/2)^2 - 0.8)``, order 0
    (reference: _kernels.py:129-162)."""

    kernel_type = "kb:beatty"

    def __init__(self, shape, grid_shape, os_grid_shape):
        if np.isscalar(shape):
            shape = (shape,)
        self.shape = tuple(int(s) for s in shape)
        grid_shape = np.atleast_1d(grid_shape)
        os_grid_shape = np.atleast_1d(os_grid_shape)
        for arr in (grid_shape, os_grid_shape):
            if arr.ndim > 1:
                raise ValueError("arr must be scalar or 1d")
            if arr.size != self.ndim:
                raise ValueError("array did not have the expected size of {}".format(self.ndim))
            if not np.all(np.mod(arr, 1) == 0):
                raise ValueError("arr contains non-integer values")
        self.grid_shape = grid_shape.astype(np.intp)
        self.os_grid_shape = os_grid_shape.astype(np.intp)
        self.alpha = [self.beatty_alpha(j, k, n) for j, k, n in
                      zip(self.shape, self.os_grid_shape, self.grid_shape)]
        self.m = [0] * self.ndim
        self.params = {}
        self._make_kernels()

    @staticmethod
    def beatty_alpha(J, K, N):
        ratio = K / N
        return np.pi * sqrt(J ** 2 / ratio ** 2 * (ratio - 0.5) ** 2 - 0.8)
