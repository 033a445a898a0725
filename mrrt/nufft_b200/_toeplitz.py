"""Toeplitz form of the normal operator ``A^H W A``.

The reference only hints at it (``return_psf`` in ``nufft_adj``, mrrt/nufft/_nufft.py:1459,
1495, 1517-1518; the operator itself lives in its external caller ``mrrt.operators``).
For ``A x[m] = sum_n x[n] exp(-i omega_m.(n - n_shift))`` the Gram operator is a
convolution,

    (A^H W A x)[n] = sum_n' T[n - n'] x[n'],     T[d] = sum_m w_m exp(+i omega_m . d),

so after ONE adjoint NUFFT of the weights onto a ``2 Nd`` image (which evaluates ``T`` on
``d in [-Nd, Nd)``) every application is: zero-pad to ``2 Nd``, FFT, multiply by the
spectrum of ``T``, inverse FFT, crop -- no pass over the ``M`` non-uniform samples at all.
The three steps run through the C ABI on a helper plan with ``Kd = 2 Nd``
(``b2n_grid_fwd`` -> ``b2n_grid_multiply`` -> ``b2n_grid_adj``).

``ToeplitzNorm`` reproduces ``A.adj(w * A.fft(x))`` up to the NUFFT approximation error of
either side (it is the closer of the two to the exact non-uniform DFT, see the test); it
is NOT bit-comparable with ``A.norm`` and ``NufftBase.norm`` never uses it silently.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._nufft import NufftBase, _ArrayKind, _TORCH_C, _f_order_memory, _prod, nufft_adj

__all__ = ["ToeplitzNorm"]


class ToeplitzNorm(object):
    """``norm(x) = A^H diag(weights) A x`` by FFT convolution on a ``2 Nd`` grid.

    Parameters
    ----------
    A : NufftBase
        The operator whose Gram is wanted (real phasing; any mode, precision, ``ortho``,
        ``n_shift``, ``adjoint_scalefactor``).
    weights : array ``(M,)``, optional
        Sample weights ``w`` (density compensation); default all ones.
    """

    def __init__(self, A, weights=None):
        if A.phasing != "real":
            raise ValueError("ToeplitzNorm needs phasing='real'")
        self.Nd, self.ndim, self.device = A.Nd, A.ndim, A.device
        self._cplx_dtype = A._cplx_dtype
        self._lib = A._lib
        self._plan = None
        cdt = _TORCH_C[A._cplx_dtype]
        N2 = tuple(2 * n for n in A.Nd)
        self.N2 = N2
        with torch.cuda.device(self.device):
            # T[d], d in [-N, N): adjoint NUFFT of the weights onto 2N with the origin at N
            A2 = NufftBase(Nd=N2, omega=A.omega, Jd=A.Jd, Kd=tuple(2 * k for k in A.Kd),
                           precision=A.precision, mode="table", Ld=A.Ld,
                           n_shift=tuple(float(n) for n in A.Nd), device=self.device)
            if weights is None:
                w = torch.ones(A.M, dtype=cdt, device=self.device)
            else:
                w = _ArrayKind(weights).to_torch(weights, self.device).to(cdt).reshape(A.M)
            psf = nufft_adj(A2, w)                                   # logical shape 2N
            del A2
            # circular layout (d mod 2N), spectrum, and the scaling of A.adj(A.fft(.))
            psf = torch.roll(psf, shifts=[-n for n in A.Nd], dims=list(range(self.ndim)))
            scale = float(A.adjoint_scalefactor) / (_prod(A.Kd) if A.ortho else 1.0)
            spec = torch.fft.fftn(psf) * scale                       # plan time, once
            self._spectrum = _f_order_memory(spec, N2).reshape(-1).contiguous()
            del psf, spec

            # helper plan: Nd -> 2Nd zero-pad / FFT / crop with unit deapodization
            plan = ctypes.c_void_p()
            arr = lambda v: (ctypes.c_int * 3)(*(list(v) + [1] * (3 - len(v))))
            _lib.check(self._lib.b2n_plan_create(
                self.ndim, arr(self.Nd), arr(N2), arr((1,) * self.ndim), 1,
                _lib.B2N_SINGLE if A.precision == "single" else _lib.B2N_DOUBLE,
                1, self.device.index, ctypes.byref(plan)))
            self._plan = plan
            sn_ptrs = (ctypes.c_void_p * 3)()
            self._ones = [np.ones(n, dtype=np.float64) for n in self.Nd]
            for d, s in enumerate(self._ones):
                sn_ptrs[d] = s.ctypes.data
            _lib.check(self._lib.b2n_plan_set_scaling(self._plan, sn_ptrs, None, 1.0,
                                                      1.0 / _prod(N2)))
            torch.cuda.current_stream(self.device).synchronize()

    @property
    def spectrum(self):
        """torch.Tensor ``2 Nd``: FFT of the point-spread function (F-ordered view)."""
        t = self._spectrum.reshape(tuple(reversed(self.N2)))
        return t.permute(*reversed(range(t.dim())))

    def __del__(self):
        plan = getattr(self, "_plan", None)
        if plan is not None and plan.value:
            try:
                self._lib.b2n_plan_destroy(plan)
            except Exception:  # pragma: no cover
                pass
            self._plan = None

    def norm(self, x):
        """``x`` of shape ``Nd`` (plus a trailing repetition axis) -> same shape."""
        kind = _ArrayKind(x)
        xt = kind.to_torch(x, self.device)
        npix = _prod(self.Nd)
        if xt.numel() == 0 or xt.numel() % npix != 0:
            raise ValueError("cannot reshape array of size {} into shape {}".format(
                xt.numel(), tuple(self.Nd) + (-1,)))
        n_reps = xt.numel() // npix
        cdt = _TORCH_C[self._cplx_dtype]
        if xt.dtype != cdt:
            xt = xt.to(cdt)
        mem = _f_order_memory(xt, self.Nd).reshape(n_reps, npix)
        grid = torch.empty((n_reps, _prod(self.N2)), dtype=cdt, device=self.device)
        out = torch.empty((n_reps,) + tuple(reversed(self.Nd)), dtype=cdt, device=self.device)
        with torch.cuda.device(self.device):
            st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(self._lib.b2n_grid_fwd(self._plan, mem.data_ptr(), grid.data_ptr(), n_reps, st))
            _lib.check(self._lib.b2n_grid_multiply(self._plan, grid.data_ptr(),
                                                   self._spectrum.data_ptr(), n_reps, st))
            _lib.check(self._lib.b2n_grid_adj(self._plan, grid.data_ptr(), out.data_ptr(), n_reps, st))
        y = out.permute(*reversed(range(out.dim())))
        if n_reps == 1:
            y = y[..., 0]
        return kind.from_torch(y)

    __call__ = norm
