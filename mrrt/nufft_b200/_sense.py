"""Coil-sensitivity (SENSE) encoding on top of ``NufftBase``.

The reference's NUFFT is called by ``mrrt.operators`` / ``mrrt.mri``'s ``MRI_Operator``
(named at mrrt/nufft/_nufft.py:3-5 and :200-201; not part of its tree), which multiplies
the image by each coil's sensitivity map before ``NufftBase.fft`` and conjugate-multiplies
and sums the coil images after ``NufftBase.adj``.  ``SenseNufft`` is that caller for the
B200 operator:

    fft(x)[:, c] = A.fft(x * smaps[..., c])
    adj(k)       = sum_c conj(smaps[..., c]) * A.adj(k[:, c])
    norm(x)      = adj(fft(x))

with the two coil steps fused into the kernels around the oversampled FFT
(``b2n_sense_fwd`` / ``b2n_sense_adj``): the ``ncoil`` coil images are never written to
HBM.  Single GPU; for several GPUs shard the coils with ``CoilShardedNufft``.
"""
import torch

from . import _lib
from ._nufft import NufftBase, _ArrayKind, _TORCH_C, _f_order_memory, _prod

__all__ = ["SenseNufft"]


class SenseNufft(object):
    """Multi-coil NUFFT encoding operator.

    Parameters
    ----------
    Nd, omega, **kwargs : as ``NufftBase`` (``order`` must stay "F").
    smaps : array ``Nd + (ncoil,)``
        Coil sensitivity maps (NumPy, torch or DLPack; cast to the precision dtype).
    """

    def __init__(self, Nd, omega, smaps, **kwargs):
        if kwargs.get("order", "F") != "F":
            raise ValueError("SenseNufft supports order='F' only")
        self.op = NufftBase(Nd=Nd, omega=omega, **kwargs)
        A = self.op
        self.Nd, self.M, self.device = A.Nd, A.M, A.device
        kind = _ArrayKind(smaps)
        s = kind.to_torch(smaps, self.device)
        if tuple(s.shape[:len(self.Nd)]) != tuple(self.Nd) or s.dim() not in (A.ndim, A.ndim + 1):
            raise ValueError("smaps must have shape Nd + (ncoil,)")
        if s.dim() == A.ndim:
            s = s[..., None]
        self.n_coils = int(s.shape[-1])
        if self.n_coils < 1:
            raise ValueError("smaps must hold at least one coil")
        cdt = _TORCH_C[A._cplx_dtype]
        # memory: first image axis fastest, coil slowest
        self._smaps = _f_order_memory(s.to(cdt), self.Nd).reshape(self.n_coils, _prod(self.Nd))

    @property
    def smaps(self):
        """torch.Tensor ``Nd + (n_coils,)``: the maps on the device (F-ordered view)."""
        t = self._smaps.reshape((self.n_coils,) + tuple(reversed(self.Nd)))
        return t.permute(*reversed(range(t.dim())))

    def _fft_dev(self, xt):
        A = self.op
        if xt.numel() != _prod(self.Nd):
            raise ValueError("cannot reshape array of size {} into shape {}".format(
                xt.numel(), tuple(self.Nd)))
        cdt = _TORCH_C[A._cplx_dtype]
        mem = _f_order_memory(xt.to(cdt).reshape(self.Nd), self.Nd).reshape(-1)
        out = torch.empty((self.n_coils, self.M), dtype=cdt, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(A._lib.b2n_sense_fwd(A._plan, mem.data_ptr(), self._smaps.data_ptr(),
                                            out.data_ptr(), self.n_coils, A._stream()))
        return out.t()                      # logical (M, n_coils), F-ordered

    def _adj_dev(self, kt):
        A = self.op
        if self.M == 0 or kt.numel() != self.M * self.n_coils:
            raise ValueError("invalid size")
        cdt = _TORCH_C[A._cplx_dtype]
        mem = _f_order_memory(kt.to(cdt).reshape(self.M, self.n_coils), (self.M,))
        out = torch.empty(tuple(reversed(self.Nd)), dtype=cdt, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(A._lib.b2n_sense_adj(A._plan, mem.data_ptr(), self._smaps.data_ptr(),
                                            out.data_ptr(), self.n_coils, A._stream()))
        return out.permute(*reversed(range(out.dim())))     # logical Nd, F-ordered

    def fft(self, x):
        """Image ``Nd`` -> samples ``(M, n_coils)``."""
        kind = _ArrayKind(x)
        return kind.from_torch(self._fft_dev(kind.to_torch(x, self.device)))

    def adj(self, k):
        """Samples ``(M, n_coils)`` -> coil-combined image ``Nd``."""
        kind = _ArrayKind(k)
        return kind.from_torch(self._adj_dev(kind.to_torch(k, self.device)))

    def norm(self, x):
        """Gram operator ``adj(fft(x))``; the samples stay on the device."""
        kind = _ArrayKind(x)
        return kind.from_torch(self._adj_dev(self._fft_dev(kind.to_torch(x, self.device))))
