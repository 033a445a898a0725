"""Non-uniform FFT operator for NVIDIA B200 (sm_100a).

``NufftBase`` keeps the constructor and the ``fft`` / ``adj`` methods of the reference
operator (mrrt/nufft/_nufft.py:122-482) so it is a drop-in for that path, and adds
``norm`` (the Gram operator ``adj(fft(x))``).  The hot path -- table interpolation
(forward gather, adjoint gridding), the sparse-matrix mode and the scale / zero-pad /
phase / crop kernels around the oversampled FFT -- runs in pre-built CUDA through the C
ABI of ``libb200nufft.so`` (include/b200nufft.h); nothing is compiled at run time and
there is no CPU fallback.

Arrays: NumPy in -> NumPy out (host<->device copies inside the call); PyTorch CUDA
tensors or any ``__dlpack__`` producer (CuPy) in -> the same kind out, zero-copy.
Multi-dimensional data is Fortran-ordered (first axis fastest, repetitions slowest) as in
the reference; an F-contiguous input is consumed without a transposition copy.
"""
import ctypes
import warnings
from math import sqrt

import numpy as np
import torch

from . import _lib
from . import _plan_math as pm
from ._kernels import BeattyKernel

__all__ = ["NufftBase", "nufft_forward", "nufft_adj"]

supported_real_types = [np.float32, np.float64]


def _as_tuple(seq, type=int, n=None):
    if np.isscalar(seq):
        if n is None:
            raise ValueError("for scalar, seq, n must be specified")
        return (type(seq),) * n
    elif n is not None and len(seq) != n:
        raise ValueError("array did not have the expected size of {}".format(n))
    return tuple(type(s) for s in seq)


def _prod(seq):
    out = 1
    for s in seq:
        out *= int(s)
    return out


_TORCH_C = {np.dtype(np.complex64): torch.complex64, np.dtype(np.complex128): torch.complex128}
_TORCH_R = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}


class _ArrayKind(object):
    """Remembers what kind of array the caller handed in, to hand the same kind back."""

    def __init__(self, x):
        self.host = False
        self.pinned = False
        if isinstance(x, torch.Tensor):
            self.kind = "torch"
            self.host = not x.is_cuda
            self.pinned = self.host and x.is_pinned()
        elif isinstance(x, np.ndarray) or np.isscalar(x) or isinstance(x, (list, tuple)):
            self.kind = "numpy"
        elif hasattr(x, "__dlpack__"):
            self.kind = "dlpack"
            self.module = type(x).__module__.split(".")[0]
        else:
            self.kind = "numpy"

    def to_torch(self, x, device):
        if self.kind == "torch":
            return x.to(device) if x.device != device else x
        if self.kind == "dlpack":
            return torch.from_dlpack(x).to(device)
        x = np.asarray(x)
        t = torch.from_numpy(np.ascontiguousarray(x) if not (x.flags.c_contiguous or x.flags.f_contiguous) else x)
        return t.to(device, non_blocking=False)

    def from_torch(self, t):
        if self.kind == "torch":
            if not self.host:
                return t
            # host tensor in -> host tensor out (pinned if the input was pinned)
            out = torch.empty_strided(t.shape, t.stride(), dtype=t.dtype, pin_memory=self.pinned)
            out.copy_(t, non_blocking=False)
            return out
        if self.kind == "dlpack":
            if self.module == "cupy":
                import cupy

                return cupy.from_dlpack(t)
            # any other DLPack producer: the result stays on the device as a torch tensor,
            # itself a DLPack producer the caller's library can consume without a copy
            return t
        return t.cpu().numpy()


def _f_order_memory(x, lead):
    """Return a C-contiguous tensor holding ``x``'s elements in Fortran order.

    ``x`` has logical shape ``lead + (reps,)``; the result has shape
    ``(reps,) + reversed(lead)``.  No copy is made if ``x`` is already F-contiguous.
    """
    return x.permute(*reversed(range(x.dim()))).contiguous()


class NufftBase(object):
    """NUFFT operator (B200).  Same parameters as the reference ``NufftBase``
    (mrrt/nufft/_nufft.py:216-233).

    Parameters
    ----------
    Nd : tuple of int
        Shape of the Cartesian grid in the spatial domain (1d, 2d or 3d).
    omega : 2d array, ``(num_samples, ndim)``
        Non-Cartesian sampling frequencies in radians (float32 or float64; NumPy,
        PyTorch or any DLPack producer).
    Jd : int or tuple, optional
        Interpolation kernel size on each axis.
    Kd : tuple, optional
        Oversampled grid size.  Default ``int(1.5 * Nd)``.
    precision : {'single', 'double', 'auto'}
    mode : {'table', 'sparse'}
    Ld : int
        Lookup-table oversampling (table length ``J * Ld + 1`` per axis).
    ortho, n_shift, phasing, adjoint_scalefactor, order : as in the reference.
    preplan_cufft, verbose : accepted for compatibility.
    on_gpu : bool
        Must be True: this operator has no CPU path.
    device : int or torch.device, optional
        CUDA device (default: the current one).
    options : dict, optional
        Integer plan options of the C ABI (``b2n_plan_set_option``).
    host_chunks : int, optional
        For HOST inputs (NumPy arrays / CPU tensors) in table mode: split the samples into
        this many contiguous ranges, each with its own plan, and pipeline the host<->device
        copies of one range against the interpolation of another (two CUDA streams).
        Results are identical; 1 disables it.
    """

    def __init__(self, Nd, omega, Jd=4, Kd=None, precision="single", mode="table",
                 Ld=1024, ortho=False, n_shift=None, phasing="real",
                 adjoint_scalefactor=1.0, preplan_cufft=True, order="F", verbose=False,
                 on_gpu=True, device=None, options=None, host_chunks=1):
        self.verbose = verbose
        self.host_chunks = max(1, int(host_chunks))
        self._children = None
        self._ctor_kwargs = dict(Jd=Jd, Kd=Kd, precision=precision, mode=mode, Ld=Ld, ortho=ortho,
                                 n_shift=n_shift, phasing=phasing,
                                 adjoint_scalefactor=adjoint_scalefactor, order=order,
                                 device=device, options=options)
        if on_gpu not in (True, False):
            raise ValueError("on_gpu must be True or False")
        if not on_gpu:
            raise ValueError(
                "mrrt.nufft_b200.NufftBase runs on the GPU only (no CPU fallback); "
                "pass on_gpu=True")
        self._plan = None
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA device required: mrrt.nufft_b200 has no CPU fallback")
        if order not in ("F", "C"):
            raise ValueError("order must be 'F' or 'C'")
        self.order = order
        self.on_gpu = True
        if np.isscalar(Nd):
            Nd = (Nd,)
        self.Nd = _as_tuple(Nd, type=int)
        self.ndim = len(self.Nd)
        if self.ndim > 3:
            raise NotImplementedError("dimensions > 3 not implemented")
        if phasing not in ("real", "complex"):
            raise ValueError(
                f"Invalid phasing: {phasing}. phasing must be 'real' or 'complex'")
        self.phasing = phasing
        self.n_mid = pm.n_mid(self.Nd, phasing)
        self.Jd = _as_tuple(Jd, type=int, n=self.ndim)
        if Kd is None:
            Kd = tuple([int(1.5 * n) for n in self.Nd])
        self.Kd = _as_tuple(Kd, type=int, n=self.ndim)
        self.ortho = ortho
        self.scale_ortho = sqrt(_prod(self.Kd)) if self.ortho else 1
        self.adjoint_scalefactor = adjoint_scalefactor
        self.preplan_cufft = preplan_cufft
        if mode not in ("table", "sparse"):
            if mode == "exact":
                raise ValueError("mode exact not implemented")
            raise ValueError("Invalid NUFFT mode: {}".format(mode))
        self.mode = mode
        self.Ld = Ld
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        elif not isinstance(device, torch.device):
            device = torch.device("cuda", int(device))
        elif device.index is None:          # torch.device("cuda"): the CURRENT device, not 0
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device

        # ---- omega: validate, remember the caller's copy for the phase computation
        kind = _ArrayKind(omega)
        if kind.kind == "numpy":
            omega_np = np.asarray(omega)
        else:
            omega_np = kind.to_torch(omega, torch.device("cpu")).numpy()
        if omega_np.ndim == 1:
            omega_np = omega_np[:, np.newaxis]
        if omega_np.shape[1] != self.ndim:
            raise ValueError("number of cols must match NUFFT dimension")
        if omega_np.dtype not in supported_real_types:
            raise ValueError(
                "omega must be one of the following types: {}".format(supported_real_types))
        if precision == "auto":
            precision = "single" if omega_np.dtype == np.float32 else "double"
        self.precision = precision
        self._real_dtype, self._cplx_dtype = pm.real_cplx_dtypes(precision)
        rdt, cdt = self._real_dtype, self._cplx_dtype
        self.M = omega_np.shape[0]
        # private copy: the sample-range sub-plans are built on the first host call and must see
        # the coordinates this plan was built from, whatever the caller does to its array
        self._omega_host = omega_np.copy() if self.host_chunks > 1 else None
        if n_shift is None:
            self.n_shift = (0.0,) * self.ndim
        else:
            self.n_shift = _as_tuple(n_shift, type=float, n=self.ndim)
        self.nargin1 = _prod(self.Nd)
        self.nargout1 = self.M

        # ---- kernel and host-side plan constants (small 1-D arrays)
        self.kernel = BeattyKernel(shape=self.Jd, grid_shape=self.Nd, os_grid_shape=self.Kd)
        self._sn1d = pm.deapodization_1d(self.Nd, self.Kd, self.Jd, self.kernel.alpha, phasing)
        omega_rdt = np.asfortranarray(omega_np.astype(rdt, copy=False))
        self._pb_angles = None
        self.phase_after = None
        self.phase_shift = None
        if phasing == "real":
            self._pb_angles = pm.phase_before_angles(self.Kd, self.n_mid, rdt)
            # the reference evaluates phase_after from omega as handed in (before the
            # precision cast) in table mode, and from the cast omega in sparse mode
            # (_nufft.py:313-315 vs :768-770)
            src = omega_rdt if mode == "sparse" else omega_np
            self.phase_after = pm.phase_after(src, self.n_mid, self.n_shift, rdt, cdt)

        # ---- device plan
        plan = ctypes.c_void_p()
        arr = lambda v: (ctypes.c_int * 3)(*(list(v) + [1] * (3 - len(v))))
        _lib.check(self._lib.b2n_plan_create(
            self.ndim, arr(self.Nd), arr(self.Kd), arr(self.Jd), int(Ld),
            _lib.B2N_SINGLE if precision == "single" else _lib.B2N_DOUBLE,
            1 if phasing == "complex" else 0, self.device.index, ctypes.byref(plan)))
        self._plan = plan
        for k, v in (options or {}).items():
            _lib.check(self._lib.b2n_plan_set_option(self._plan, k.encode(), int(v)))

        with torch.cuda.device(self.device):
            stream = self._stream()
            self.omega = torch.from_numpy(omega_rdt.T.copy()).to(self.device)  # [ndim, M]
            _lib.check(self._lib.b2n_plan_set_points(
                self._plan, self.omega.data_ptr(), self.M, _lib.B2N_COORD_OMEGA, stream))
            self.omega = self.omega.t()  # logical (M, ndim), F-ordered like the reference

            sample_phase = None
            if mode == "sparse":
                self._init_sparsemat(omega_rdt)
                _lib.check(self._lib.b2n_plan_set_option(self._plan, b"sparse_mode", 1))
                if phasing == "real":
                    sample_phase = self.phase_after
            else:
                odd_L = Ld % 2 == 1
                odd_J = np.mod(self.Jd, 2) == 1
                if odd_L and any(odd_J):
                    warnings.warn("accuracy may be compromised when L and J are both odd")
                self._init_table()
                if phasing == "real":
                    sample_phase = self.phase_after
                elif any(s != 0 for s in self.n_shift):
                    # _nufft.py:898-901
                    self.phase_shift = np.exp(
                        1j * np.dot(omega_rdt, np.asarray(self.n_shift)))
                    sample_phase = self.phase_shift.astype(cdt)
            self._sample_phase_dev = None
            if sample_phase is not None:
                ph = torch.from_numpy(np.ascontiguousarray(sample_phase.astype(cdt, copy=False)))
                self._sample_phase_dev = ph.to(self.device)
                _lib.check(self._lib.b2n_plan_set_sample_phase(
                    self._plan, self._sample_phase_dev.data_ptr(), stream))

            # scaling constants for the full transforms
            sn_ptrs = (ctypes.c_void_p * 3)()
            self._sn1d_c = [np.ascontiguousarray(s, dtype=np.float64) for s in self._sn1d]
            for d, s in enumerate(self._sn1d_c):
                sn_ptrs[d] = s.ctypes.data
            pb_ptrs = None
            if self._pb_angles is not None:
                pb_ptrs = (ctypes.c_void_p * 3)()
                self._pb_c = [np.ascontiguousarray(a, dtype=rdt) for a in self._pb_angles]
                for d, a in enumerate(self._pb_c):
                    pb_ptrs[d] = a.ctypes.data
            fwd_scale = 1.0 / self.scale_ortho if self.ortho else 1.0
            adj_scale = (self.adjoint_scalefactor / self.scale_ortho if self.ortho
                         else float(self.adjoint_scalefactor))
            _lib.check(self._lib.b2n_plan_set_scaling(
                self._plan, sn_ptrs, pb_ptrs, fwd_scale, adj_scale))
            torch.cuda.current_stream().synchronize()

    # ------------------------------------------------------------------ plan pieces
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _init_table(self):
        """Per-axis lookup tables (_nufft.py:880-935)."""
        rdt, cdt = self._real_dtype, self._cplx_dtype
        self.h = []
        for d in range(self.ndim):
            h = pm.lookup_table(self.Nd[d], self.Jd[d], self.Kd[d], self.Ld, self.phasing)
            self.h.append(np.ascontiguousarray(h.astype(cdt if self.phasing == "complex" else rdt)))
        ptrs = (ctypes.c_void_p * 3)()
        for d, h in enumerate(self.h):
            ptrs[d] = h.ctypes.data
        _lib.check(self._lib.b2n_plan_set_tables(self._plan, ptrs))

    def _init_sparsemat(self, omega_rdt):
        """Interpolation matrix in fixed-width row form (_nufft.py:751-877).

        Per-axis coefficients are evaluated on the host exactly like the reference
        (KB at float64 arguments); the products over axes, the conjugation, the
        ``n_shift`` phase of complex phasing and the cast are done on the device.
        """
        coef_dev, kidx_dev = [], []
        coef_ptrs = (ctypes.c_void_p * 3)()
        kidx_ptrs = (ctypes.c_void_p * 3)()
        for d in range(self.ndim):
            u, kd = pm.axis_coefficients(omega_rdt[:, d], self.Nd[d], self.Jd[d], self.Kd[d],
                                         self.kernel.alpha[d], self.phasing)
            u = np.ascontiguousarray(u.T)           # [M, J], tap fastest
            kd = np.ascontiguousarray(kd.T.astype(np.int32))
            coef_dev.append(torch.from_numpy(u).to(self.device))
            kidx_dev.append(torch.from_numpy(kd).to(self.device))
            coef_ptrs[d] = coef_dev[-1].data_ptr()
            kidx_ptrs[d] = kidx_dev[-1].data_ptr()
        row_phase = None
        rp_ptr = None
        if self.phasing == "complex" and any(s != 0 for s in self.n_shift):
            ph = np.exp(1j * np.dot(omega_rdt, np.asarray(self.n_shift)))
            row_phase = torch.from_numpy(np.ascontiguousarray(ph.astype(np.complex128))).to(self.device)
            rp_ptr = row_phase.data_ptr()
        _lib.check(self._lib.b2n_plan_set_sparse(
            self._plan, coef_ptrs, kidx_ptrs, self.M, rp_ptr, self._stream()))
        torch.cuda.current_stream(self.device).synchronize()
        self.sparse_format = "ELL"

    # ------------------------------------------------------------------ attributes
    @property
    def sn(self):
        """ndarray: dense deapodization array (built on demand; the device only keeps
        its separable 1-D factors)."""
        return pm.dense_sn(self._sn1d, self.Nd).astype(self._real_dtype)

    @property
    def phase_before(self):
        """ndarray or None: dense FFT-shift phase on the Kd grid (built on demand)."""
        if self._pb_angles is None:
            return None
        return pm.dense_phase_before(self._pb_angles, self._cplx_dtype)

    @property
    def tm(self):
        """torch.Tensor (M, ndim): sample coordinates in grid units, ``omega/(2 pi/K)``
        evaluated in the precision dtype (_nufft.py:338-342)."""
        buf = torch.empty((self.ndim, self.M), dtype=_TORCH_R[self._real_dtype], device=self.device)
        _lib.check(self._lib.b2n_plan_get_points(self._plan, buf.data_ptr(), None, None, None,
                                                 self._stream()))
        return buf.t()

    def bin_sort(self):
        """(bin_ids int32 [M], keys int64 [M], perm int32 [M]) as device tensors."""
        bins = torch.empty(self.M, dtype=torch.int32, device=self.device)
        keys = torch.empty(self.M, dtype=torch.int64, device=self.device)
        perm = torch.empty(self.M, dtype=torch.int32, device=self.device)
        _lib.check(self._lib.b2n_plan_get_points(self._plan, None, bins.data_ptr(), keys.data_ptr(),
                                                 perm.data_ptr(), self._stream()))
        return bins, keys, perm

    def forward_slots(self):
        """uint32 slot list of the paired forward kernel as an int64 device tensor
        ((sorted position << 1) | has_partner); empty when the plan has none."""
        n = int(self._lib.b2n_plan_num_slots(self._plan))
        raw = torch.empty(max(n, 0), dtype=torch.int32, device=self.device)
        if n > 0:
            _lib.check(self._lib.b2n_plan_get_slots(self._plan, raw.data_ptr(), self._stream()))
        return raw.to(torch.int64) & 0xFFFFFFFF

    @property
    def tile(self):
        return tuple(int(self._lib.b2n_plan_get_option(self._plan, ("tile%d" % (d + 1)).encode()))
                     for d in range(self.ndim))

    @property
    def p(self):
        """scipy.sparse.csr_matrix or None: host copy of the interpolation matrix
        (sparse mode), for inspection."""
        if self.mode != "sparse":
            return None
        import scipy.sparse

        nnz = int(self._lib.b2n_plan_sparse_nnz(self._plan))
        nnzr = nnz // max(self.M, 1)
        vdt = self._cplx_dtype if self.phasing == "complex" else self._real_dtype
        tdt = _TORCH_C[vdt] if self.phasing == "complex" else _TORCH_R[vdt]
        vals = torch.empty(nnz, dtype=tdt, device=self.device)
        cols = torch.empty(nnz, dtype=torch.int32, device=self.device)
        _lib.check(self._lib.b2n_plan_get_sparse(self._plan, vals.data_ptr(), cols.data_ptr(),
                                                 self._stream()))
        indptr = np.arange(0, nnz + 1, nnzr, dtype=np.int64) if nnzr else np.zeros(self.M + 1, np.int64)
        m = scipy.sparse.csr_matrix((vals.cpu().numpy(), cols.cpu().numpy(), indptr),
                                    shape=(self.M, _prod(self.Kd)))
        m.sum_duplicates()
        return m

    def option(self, name):
        return int(self._lib.b2n_plan_get_option(self._plan, name.encode()))

    def set_option(self, name, value):
        _lib.check(self._lib.b2n_plan_set_option(self._plan, name.encode(), int(value)))

    def kernel_timing(self):
        """(fwd_ms_total, fwd_launches, adj_ms_total, adj_launches) of the interpolation
        kernels since the last call; needs ``set_option("profile", 1)``."""
        out = (ctypes.c_double * 4)()
        _lib.check(self._lib.b2n_plan_get_timing(self._plan, out))
        return tuple(out)

    @property
    def launch_count(self):
        return int(self._lib.b2n_plan_launch_count(self._plan))

    @property
    def device_bytes(self):
        return int(self._lib.b2n_plan_device_bytes(self._plan))

    def __del__(self):
        plan = getattr(self, "_plan", None)
        if plan is not None and plan.value:
            try:
                self._lib.b2n_plan_destroy(plan)
            except Exception:  # pragma: no cover
                pass
            self._plan = None

    # ------------------------------------------------------------------ transforms
    def _swap_reps(self, x, narg):
        if x.numel() != narg:
            x = x.permute(*(tuple(range(1, x.dim())) + (0,)))
        return x

    def _unswap_reps(self, x, narg):
        if x.numel() != narg:
            x = x.permute(*((x.dim() - 1,) + tuple(range(x.dim() - 1))))
        return x

    # ------------------------------------------------------------------ host pipeline
    def _pipelined(self, kind):
        return (self.host_chunks > 1 and self.mode == "table" and self.M >= self.host_chunks
                and (kind.kind == "numpy" or kind.host))

    def _ensure_children(self):
        """Sub-operators over contiguous sample ranges (built on first host call)."""
        if self._children is None:
            from ._sharded import shard_range

            kw = dict(self._ctor_kwargs)
            kw["device"] = self.device
            self._children = []
            for k in range(self.host_chunks):
                lo, hi = shard_range(self.M, self.host_chunks, k)
                self._children.append((lo, hi, NufftBase(self.Nd, self._omega_host[lo:hi], **kw)))
            self._omega_host = None              # the sub-plans hold what they need
            # one stream per copy direction: the two DMA directions of the link are
            # independent, so the samples of one call can travel back while the inputs of
            # the next call come in
            self._h2d_stream = torch.cuda.Stream(device=self.device)
            self._d2h_stream = torch.cuda.Stream(device=self.device)
        return self._children

    def synchronize(self):
        """Wait for every transform issued with ``non_blocking=True`` (and everything else
        on this operator's streams); their host results are valid afterwards."""
        with torch.cuda.device(self.device):
            torch.cuda.current_stream(self.device).synchronize()
            if self._children is not None:
                self._h2d_stream.synchronize()
                self._d2h_stream.synchronize()

    def _host_tensor(self, x):
        t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.asarray(x))
        cdt = _TORCH_C[self._cplx_dtype]
        return t if t.dtype == cdt else t.to(cdt)

    def _fft_host(self, x, kind, non_blocking=False):
        """fft for a HOST array: H2D image, one spectrum, then per sample range the
        interpolation on the compute stream while the previous range's samples travel
        back on the D2H stream.

        Buffer lifetimes follow the caching allocator's stream rules: a buffer written by a
        copy stream is allocated under that stream and ``record_stream``-ed for the compute
        stream (and vice versa), so no host synchronisation is needed to keep it alive and
        ``non_blocking=True`` can return while the copies are still in flight."""
        xt = self._host_tensor(x)
        if self.order == "C":
            xt = self._swap_reps(xt, self.nargin1)
        if xt.numel() == 0 or xt.numel() % self.nargin1 != 0:
            print("Input signal has the wrong size.")
            raise ValueError("cannot reshape array of size {} into shape {}".format(
                xt.numel(), tuple(self.Nd) + (-1,)))
        n_reps = xt.numel() // self.nargin1
        children = self._ensure_children()
        cdt = _TORCH_C[self._cplx_dtype]
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream(self.device)
            h2d, d2h = self._h2d_stream, self._d2h_stream
            mem_h = _f_order_memory(xt, self.Nd).reshape(n_reps, self.nargin1)
            with torch.cuda.stream(h2d):
                mem = torch.empty(mem_h.shape, dtype=cdt, device=self.device)
                mem.copy_(mem_h, non_blocking=True)
            mem.record_stream(main)
            main.wait_stream(h2d)
            grid = torch.empty((n_reps, _prod(self.Kd)), dtype=cdt, device=self.device)
            _lib.check(self._lib.b2n_grid_fwd(self._plan, mem.data_ptr(), grid.data_ptr(), n_reps,
                                              self._stream()))
            out_h = torch.empty((n_reps, self.M), dtype=cdt, pin_memory=True)
            for lo, hi, ch in children:
                out_k = torch.empty((n_reps, hi - lo), dtype=cdt, device=self.device)
                _lib.check(self._lib.b2n_interp_fwd(ch._plan, grid.data_ptr(), out_k.data_ptr(),
                                                    n_reps, 1, self._stream()))
                ev = torch.cuda.Event()
                ev.record(main)
                out_k.record_stream(d2h)
                with torch.cuda.stream(d2h):
                    d2h.wait_event(ev)
                    out_h[:, lo:hi].copy_(out_k, non_blocking=True)
            if not non_blocking:
                d2h.synchronize()
        out = out_h.t()
        if n_reps == 1:
            out = out[..., 0]
        if self.order == "C":
            out = self._unswap_reps(out, self.nargout1)
        return out.numpy() if kind.kind == "numpy" else out

    def _adj_host(self, k, kind, non_blocking=False):
        """adj for a HOST array: sample ranges are copied in on the H2D stream while the
        previous range is gridded (accumulating into one grid) on the compute stream; the
        image travels back on the D2H stream."""
        kt = self._host_tensor(k)
        if self.order == "C":
            kt = self._swap_reps(kt, self.nargout1)
        if self.M == 0 or kt.numel() == 0 or kt.numel() % self.M != 0:
            raise ValueError("invalid size")
        n_reps = kt.numel() // self.M
        children = self._ensure_children()
        cdt = _TORCH_C[self._cplx_dtype]
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream(self.device)
            h2d, d2h = self._h2d_stream, self._d2h_stream
            mem_h = _f_order_memory(kt, (self.M,)).reshape(n_reps, self.M)
            grid = torch.empty((n_reps, _prod(self.Kd)), dtype=cdt, device=self.device)
            for idx, (lo, hi, ch) in enumerate(children):
                ev = torch.cuda.Event()
                with torch.cuda.stream(h2d):
                    buf = torch.empty((n_reps, hi - lo), dtype=cdt, device=self.device)
                    buf.copy_(mem_h[:, lo:hi], non_blocking=True)
                    ev.record(h2d)
                buf.record_stream(main)
                main.wait_event(ev)
                # bit 0: apply the sample phase, bit 1: accumulate into the grid
                _lib.check(self._lib.b2n_interp_adj(ch._plan, buf.data_ptr(), grid.data_ptr(), n_reps,
                                                    1 | (2 if idx > 0 else 0), self._stream()))
            out = torch.empty((n_reps,) + tuple(reversed(self.Nd)), dtype=cdt, device=self.device)
            _lib.check(self._lib.b2n_grid_adj(self._plan, grid.data_ptr(), out.data_ptr(), n_reps,
                                              self._stream()))
            out_h = torch.empty(out.shape, dtype=cdt, pin_memory=True)
            ev = torch.cuda.Event()
            ev.record(main)
            out.record_stream(d2h)
            with torch.cuda.stream(d2h):
                d2h.wait_event(ev)
                out_h.copy_(out, non_blocking=True)
            if not non_blocking:
                d2h.synchronize()
        x = out_h.permute(*reversed(range(out_h.dim())))
        if n_reps == 1:
            x = x[..., 0]
        if self.order == "C":
            x = self._unswap_reps(x, self.nargin1)
        return x.numpy() if kind.kind == "numpy" else x

    def fft(self, x, non_blocking=False):
        """Forward NUFFT (uniform spatial -> non-uniform frequency).

        ``x`` has shape ``Nd`` (plus a trailing repetition axis for ``order="F"``, a
        leading one for ``order="C"``).  Returns ``(M,)`` / ``(M, reps)`` / ``(reps, M)``.
        Reference: _nufft.py:425-452.

        ``non_blocking=True`` (pinned HOST torch tensors with ``host_chunks > 1`` only; as
        in ``torch.Tensor.to``): the call returns a pinned host tensor while its copies are
        still in flight, so the host<->device traffic of consecutive calls overlaps in both
        link directions; the result is valid after ``synchronize()``.
        """
        kind = _ArrayKind(x)
        if self._pipelined(kind):
            return self._fft_host(x, kind, non_blocking and kind.kind == "torch" and kind.pinned)
        xt = kind.to_torch(x, self.device)
        if self.order == "C":
            xt = self._swap_reps(xt, self.nargin1)
        k = nufft_forward(self, xt)
        if self.order == "C":
            k = self._unswap_reps(k, self.nargout1)
        return kind.from_torch(k)

    def adj(self, k, non_blocking=False):
        """Adjoint NUFFT (non-uniform frequency -> uniform spatial).
        Reference: _nufft.py:454-482.  ``non_blocking``: see ``fft``."""
        kind = _ArrayKind(k)
        if self._pipelined(kind):
            return self._adj_host(k, kind, non_blocking and kind.kind == "torch" and kind.pinned)
        kt = kind.to_torch(k, self.device)
        if self.order == "C":
            kt = self._swap_reps(kt, self.nargout1)
        x = nufft_adj(self, kt)
        if self.order == "C":
            x = self._unswap_reps(x, self.nargin1)
        return kind.from_torch(x)

    def norm(self, x):
        """Gram (normal) operator ``adj(fft(x))``.  Not present upstream (the reference
        class has no ``norm``); defined as the composition of its two transforms."""
        kind = _ArrayKind(x)
        xt = kind.to_torch(x, self.device)
        if self.order == "C":
            xt = self._swap_reps(xt, self.nargin1)
        out = nufft_adj(self, nufft_forward(self, xt))
        if self.order == "C":
            out = self._unswap_reps(out, self.nargin1)
        return kind.from_torch(out)

    def __str__(self):
        keys = ["Nd", "Kd", "Jd", "Ld", "M", "ndim", "mode", "precision", "phasing", "order",
                "ortho", "n_shift", "adjoint_scalefactor", "device"]
        return "".join("{} = {}\n".format(k, getattr(self, k)) for k in keys)


def _complexify(obj, x):
    cdt = _TORCH_C[obj._cplx_dtype]
    if x.dtype != cdt:
        x = x.to(cdt)
    return x


def nufft_forward(obj, x, copy_x=True, grid_only=False, xp=None):
    """Forward NUFFT driver (reference: _nufft.py:1275-1397).

    ``x`` is a device tensor of logical shape ``Nd + (reps,)`` (or ``(prod(Kd), reps)``
    with ``grid_only=True``, which runs the interpolation stage alone).
    """
    Nd, Kd = obj.Nd, obj.Kd
    if not isinstance(x, torch.Tensor):
        x = _ArrayKind(x).to_torch(x, obj.device)
    lead = (_prod(Kd),) if grid_only else tuple(Nd)
    nlead = _prod(lead)
    if x.numel() == 0 or x.numel() % nlead != 0:
        print("Input signal has the wrong size.")
        raise ValueError("cannot reshape array of size {} into shape {}".format(
            x.numel(), tuple(lead) + (-1,)))
    n_reps = x.numel() // nlead
    x = _complexify(obj, x)
    # memory layout: first axis fastest, repetitions slowest
    mem = _f_order_memory(x, lead).reshape(n_reps, nlead)
    out = torch.empty((n_reps, obj.M), dtype=mem.dtype, device=obj.device)
    with torch.cuda.device(obj.device):
        stream = obj._stream()
        if grid_only:
            if obj.mode == "sparse":
                rc = obj._lib.b2n_spmv_fwd(obj._plan, mem.data_ptr(), out.data_ptr(), n_reps, 0, stream)
            else:
                # phase_shift of complex phasing belongs to the interpolation stage
                # (_nufft.py:1086-1095); phase_after does not
                rc = obj._lib.b2n_interp_fwd(obj._plan, mem.data_ptr(), out.data_ptr(), n_reps,
                                             1 if obj.phase_shift is not None else 0, stream)
        else:
            rc = obj._lib.b2n_nufft_fwd(obj._plan, mem.data_ptr(), out.data_ptr(), n_reps, stream)
    _lib.check(rc)
    out = out.t()  # logical (M, reps), F-ordered
    if grid_only:
        return out
    if n_reps == 1:
        out = out[..., 0]
    return out


def nufft_adj(obj, xk, copy=True, return_psf=False, grid_only=False, xp=None):
    """Adjoint NUFFT driver (reference: _nufft.py:1459-1578).

    ``grid_only`` returns the gridded samples ``(prod(Kd), reps)`` (conj(phase_after)
    applied first, as upstream).  ``return_psf`` (upstream "EXPERIMENTAL",
    _nufft.py:1495,1517-1518) grids WITHOUT conj(phase_after) and returns the first
    repetition with shape ``Kd``."""
    Nd, Kd = obj.Nd, obj.Kd
    if not isinstance(xk, torch.Tensor):
        xk = _ArrayKind(xk).to_torch(xk, obj.device)
    if obj.M == 0 or xk.numel() % obj.M != 0 or xk.numel() == 0:
        raise ValueError("invalid size")
    n_reps = xk.numel() // obj.M
    xk = _complexify(obj, xk)
    mem = _f_order_memory(xk, (obj.M,)).reshape(n_reps, obj.M)
    with torch.cuda.device(obj.device):
        stream = obj._stream()
        if grid_only or return_psf:
            if return_psf and not grid_only:
                mem, n_reps = mem[:1], 1            # only the first repetition is returned
            out = torch.empty((n_reps, _prod(Kd)), dtype=mem.dtype, device=obj.device)
            # the reference multiplies by conj(phase_after) BEFORE the gridding stage, so
            # its grid_only adjoint includes it (_nufft.py:1495-1509), unlike grid_only forward;
            # return_psf skips it, but the phase_shift of complex phasing belongs to the
            # table interpolator itself (_nufft.py:1141-1150) and stays
            if obj.mode == "sparse":
                rc = obj._lib.b2n_spmv_adj(obj._plan, mem.data_ptr(), out.data_ptr(), n_reps,
                                           0 if return_psf else 1, stream)
            else:
                ph = (1 if obj.phase_shift is not None else 0) if return_psf else 1
                rc = obj._lib.b2n_interp_adj(obj._plan, mem.data_ptr(), out.data_ptr(), n_reps,
                                             ph, stream)
            _lib.check(rc)
            if grid_only:
                return out.t()
            return out[0].reshape(tuple(reversed(Kd))).permute(*reversed(range(len(Kd))))
        out = torch.empty((n_reps,) + tuple(reversed(Nd)), dtype=mem.dtype, device=obj.device)
        rc = obj._lib.b2n_nufft_adj(obj._plan, mem.data_ptr(), out.data_ptr(), n_reps, stream)
    _lib.check(rc)
    x = out.permute(*reversed(range(out.dim())))  # logical Nd + (reps,)
    if n_reps == 1:
        x = x[..., 0]
    return x
