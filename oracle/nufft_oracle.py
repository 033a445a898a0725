"""CPU oracle for the NUFFT hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``mrrt.nufft_b200``) never does: it has no CPU fallback.

What is here
------------
A NumPy restatement of the reference's plan construction and transform drivers,
each function citing the reference ``file:line`` it follows (paths relative to
``/root/reference/mrrt/nufft``), plus ctypes bindings to two C interpolator
libraries:

* ``engine="port"``       ``oracle/liboracle_interp.so`` -- our plain-C restatement
                          (``oracle/interp_oracle_impl.h``).
* ``engine="reference"``  ``oracle/_ref/libnufft_table_ref.so`` -- the reference's
                          own ``_extensions/c/nufft_table.c`` compiled unmodified
                          with its own flags (``-O2 -ffast-math -fopenmp``).

Parity pinning: ``tests/test_oracle.py`` checks this module against the golden
vectors in ``tests/golden/`` that ``tests/golden/make_golden.py`` produced by
running the real reference package (``NufftBase``) in this container, and checks
the "port" engine against the "reference" engine.

The third-party dependency ``mrrt.utils`` (github.com/mritools/mrrt.utils, version
unpinned by the reference: requirements/default.txt:4) is not vendored by the
reference; the only hot-path arithmetic it carries is ``fftn``/``ifftn``, restated
here as ``numpy.fft.fftn/ifftn(x, s=Kd, axes=...)`` (zero-pad at the end,
unnormalised forward, 1/n inverse), which the reference's own dtft-based tests pin
(tests/test_nufft.py:99-324).
"""
import ctypes
import os
from math import sqrt

import numpy as np
import scipy.sparse
from scipy.special import i0, jv

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT_LIB = os.path.join(_HERE, "liboracle_interp.so")
_REF_LIB = os.path.join(_HERE, "_ref", "libnufft_table_ref.so")

_libs = {}


def have_reference_engine():
    return os.path.exists(_REF_LIB)


def _lib(engine):
    if engine not in _libs:
        path = {"port": _PORT_LIB, "reference": _REF_LIB}[engine]
        if not os.path.exists(path):
            raise RuntimeError(
                "oracle library %s missing: run `make -C oracle`" % path
            )
        _libs[engine] = ctypes.CDLL(path)
    return _libs[engine]


# ----------------------------------------------------------------------------
# kernel mathematics
# ----------------------------------------------------------------------------
def beatty_alpha(J, K, N):
    """Beatty et al. eq. 5 shape parameter (_kernels.py:151-154)."""
    k_n = K / N
    return np.pi * sqrt(J ** 2 / k_n ** 2 * (k_n - 0.5) ** 2 - 0.8)


def kaiser_bessel(x, J, alpha):
    """KB window, order m=0 (_kaiser_bessel.py:133-149)."""
    x = np.asarray(x)
    ii = (2 * np.abs(x) < J).nonzero()
    tmp = 2 * x[ii] / J
    tmp *= tmp
    f = np.sqrt(1 - tmp)
    kb = np.zeros_like(x)
    kb[ii] = i0(alpha * f) / float(i0(alpha))
    return kb


def kaiser_bessel_ft(u, J, alpha):
    """Fourier transform of the KB window, m=0, d=1 (_kaiser_bessel.py:197-226)."""
    u = np.asarray(u, dtype=np.float64)
    tmp = (np.pi * J) * u
    tmp *= tmp
    tmp -= alpha * alpha
    z = np.lib.scimath.sqrt(tmp)
    nu = 0.5
    const1 = (2 * np.pi) ** 0.5 * (J / 2.0) / i0(alpha)
    y = const1 * jv(nu, z)
    y = y / z ** nu
    return np.real(y)


def nufft_offset(om, J, K):
    """Window origin ``floor(om/gam - J/2)`` (_utils.py:20-43)."""
    om = np.asanyarray(om)
    gam = 2 * np.pi / K
    return np.floor(om / gam - J / 2.0)


def nufft_coef(om, J, K, alpha):
    """Per-sample kernel arguments and KB coefficients [J, M] (_utils.py:47-86)."""
    om = np.atleast_1d(np.squeeze(om))
    gam = 2 * np.pi / K
    dk = om / gam - nufft_offset(om, J, K)
    arg = -np.arange(1, J + 1)[:, None] + dk[None, :]
    return kaiser_bessel(arg, J, alpha), arg


def n_mid_of(Nd, phasing):
    """_nufft.py:623-628."""
    if phasing == "real":
        return tuple(n // 2 for n in Nd)
    return tuple((n - 1) / 2.0 for n in Nd)


def scaling_factors_1d(Nd, Kd, Jd, phasing):
    """Per-axis deapodization vectors (_nufft.py:737-746)."""
    n_mid = n_mid_of(Nd, phasing)
    out = []
    for d in range(len(Nd)):
        start = -n_mid[d]
        nc = np.arange(start, start + Nd[d])
        alpha = beatty_alpha(Jd[d], Kd[d], Nd[d])
        out.append(1 / kaiser_bessel_ft(nc / Kd[d], Jd[d], alpha))
    return out


def scaling_factors(Nd, Kd, Jd, phasing):
    """Dense ``sn`` (_nufft.py:727-748): outer product, float64."""
    sn = np.array([1.0])
    for tmp in scaling_factors_1d(Nd, Kd, Jd, phasing):
        sn = np.outer(sn.ravel(), tmp.conj())
    return sn.reshape(Nd)


def phase_before(Kd, n_mid, rdt, cdt):
    """_nufft.py:703-715 (arithmetic in the precision real dtype)."""
    ndim = len(Kd)
    phase = (2 * np.pi / Kd[0] * n_mid[0]) * np.arange(Kd[0], dtype=rdt)
    for d in range(1, ndim):
        tmp = (2 * np.pi / Kd[d] * n_mid[d]) * np.arange(Kd[d], dtype=rdt)
        phase = phase.reshape((phase.shape) + (1,)) + tmp.reshape(
            (1,) * d + (tmp.size,)
        )
    return np.exp(1j * phase).astype(cdt, copy=False)


def phase_after_angle(omega, n_mid, n_shift, rdt):
    """Argument of ``phase_after`` (_nufft.py:717-723)."""
    shift_vec = [(s - m) for s, m in zip(n_shift, n_mid)]
    return np.dot(omega, np.asarray(shift_vec, dtype=rdt))


def phase_after(omega, n_mid, n_shift, rdt, cdt):
    """_nufft.py:717-724."""
    phase = np.exp(1j * phase_after_angle(omega, n_mid, n_shift, rdt))
    return phase.astype(cdt, copy=False)


# ----------------------------------------------------------------------------
# sparse matrix and lookup table
# ----------------------------------------------------------------------------
def sparse_matrix(omega, Nd, Jd, Kd, phasing, n_shift, rdt, cdt):
    """Interpolation matrix P [M, prod(Kd)] in CSC (_nufft.py:751-877).

    ``omega`` must already be in the precision real dtype ``rdt``.
    """
    ndim = len(Nd)
    M = omega.shape[0]
    ud = {}
    kd = {}
    for d in range(ndim):
        N, J, K = Nd[d], Jd[d], Kd[d]
        alpha = beatty_alpha(J, K, N)
        c, arg = nufft_coef(omega[:, d], J, K, alpha)
        koff = nufft_offset(omega[:, d], J, K)
        kd[d] = np.mod(np.arange(1, J + 1)[:, None] + koff[None, :], K)
        if phasing == "complex":
            gam = 2 * np.pi / K
            phase = np.exp((1j * gam * (N - 1) / 2.0) * arg)
        else:
            phase = 1.0
        ud[d] = phase * c
    kk = kd[0]
    uu = ud[0]
    for d in range(1, ndim):
        Jprod = int(np.prod(Jd[: d + 1]))
        tmp = kd[d] * int(np.prod(Kd[:d]))
        kk = (kk[:, None, :] + tmp[None, :, :]).reshape((Jprod, M), order="F")
        uu = (uu[:, None, :] * ud[d][None, :, :]).reshape((Jprod, M), order="F")
    if np.iscomplexobj(uu):
        uu = uu.conj()
    if phasing == "complex":
        if any(s != 0 for s in n_shift):
            ph = np.exp(1j * np.dot(omega, np.asarray(n_shift)))
            uu = uu * ph.reshape((1, -1), order="F")
        sparse_dtype = cdt
    else:
        sparse_dtype = rdt
    mm = np.tile(np.arange(M), (int(np.prod(Jd)), 1))
    p = scipy.sparse.coo_matrix(
        (uu.ravel(order="F"), (mm.ravel(order="F"), kk.ravel(order="F"))),
        shape=(M, int(np.prod(Kd))),
        dtype=sparse_dtype,
    )
    return p.tocsc()


def make_table(N, J, K, L, phasing):
    """One axis of the lookup table, ``how="fast"`` (_nufft.py:1195-1243).

    The reference builds a dummy 1-D *single precision* sparse operator
    (precision defaults to "single", _nufft.py:222) and reads J columns of its
    matrix; so the table is float32-accurate in every precision.
    """
    if N % 2 == 0:
        t1 = J / 2.0 - 1 + np.arange(L) / L
    else:
        t1 = J / 2.0 - 1 + np.arange(1, L + 1) / L
    omega1 = (t1 * 2 * np.pi / K).astype(np.float32)[:, None]
    p = sparse_matrix(
        omega1, (N,), (J,), (K,), phasing, (0.0,), np.float32, np.complex64
    )
    h = np.asarray(p[:, np.arange(J - 1, -1, -1)].todense()).ravel(order="F")
    if N % 2 == 0:
        h = np.concatenate((h, np.atleast_1d(h[0])), axis=0)
    else:
        h = np.concatenate((np.atleast_1d(h[-1]), h), axis=0)
    return h


# ----------------------------------------------------------------------------
# C interpolators (ctypes)
# ----------------------------------------------------------------------------
def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def set_table_order(order):
    """Table lookup order of the C restatement ("port" engine only): 1 linear interpolation
    (default, the reference's CPU behaviour), 0 the entry at floor(p) (the reference's GPU
    templates with ``order=0``, cuda/cupy.py:96-98)."""
    lib = _lib("port")
    for pre in ("orc_f32_", "orc_f64_"):
        getattr(lib, pre + "set_table_order")(ctypes.c_int(int(order)))


def _interp_port(direction, Kd, Jd, L, h, tm, data):
    """Call our C restatement. data: [prod(Kd)] (fwd) or [M] (adj), complex."""
    ndim = len(Kd)
    rdt = tm.dtype
    cdt = np.complex64 if rdt == np.float32 else np.complex128
    pre = "orc_f32_" if rdt == np.float32 else "orc_f64_"
    lib = _lib("port")
    h_cplx = int(np.iscomplexobj(h[0]))
    hs = [np.ascontiguousarray(x, dtype=cdt if h_cplx else rdt) for x in h]
    while len(hs) < 3:
        hs.append(hs[0])
    K = (ctypes.c_int * 3)(*(list(Kd) + [1] * (3 - ndim)))
    J = (ctypes.c_int * 3)(*(list(Jd) + [1] * (3 - ndim)))
    tm = np.asfortranarray(tm)
    M = tm.shape[0]
    data = np.ascontiguousarray(data, dtype=cdt)
    if direction == "fwd":
        out = np.empty(M, dtype=cdt)
        fn = getattr(lib, pre + "interp_fwd")
    else:
        out = np.empty(int(np.prod(Kd)), dtype=cdt)
        fn = getattr(lib, pre + "interp_adj")
    fn.restype = None
    fn(ctypes.c_int(ndim), K, J, ctypes.c_int(L), _ptr(hs[0]), _ptr(hs[1]),
       _ptr(hs[2]), ctypes.c_int(h_cplx), _ptr(tm), ctypes.c_long(M),
       _ptr(data), _ptr(out))
    return out


def _interp_reference(direction, Kd, Jd, L, h, tm, data):
    """Call the reference's own compiled C (planar re/im arrays, as its Cython
    shim does: _nufft_table.pyx:52-61,144-152,357-367)."""
    ndim = len(Kd)
    rdt = tm.dtype
    cdt = np.complex64 if rdt == np.float32 else np.complex128
    lib = _lib("reference")
    h_cplx = np.iscomplexobj(h[0])
    name = "%s_interp%d_table1_%s_%s" % (
        "float" if rdt == np.float32 else "double",
        ndim,
        "complex" if h_cplx else "real",
        "forward" if direction == "fwd" else "adj",
    )
    fn = getattr(lib, name)
    fn.restype = None
    tm = np.asfortranarray(tm)
    M = tm.shape[0]
    data = np.asarray(data, dtype=cdt)
    d_r = np.ascontiguousarray(data.real)
    d_i = np.ascontiguousarray(data.imag)
    n_out = M if direction == "fwd" else int(np.prod(Kd))
    o_r = np.zeros(n_out, dtype=rdt)
    o_i = np.zeros(n_out, dtype=rdt)
    keep = []
    hargs = []
    for x in h:
        if h_cplx:
            xr = np.ascontiguousarray(x.real, dtype=rdt)
            xi = np.ascontiguousarray(x.imag, dtype=rdt)
            keep += [xr, xi]
            hargs += [_ptr(xr), _ptr(xi)]
        else:
            xr = np.ascontiguousarray(x, dtype=rdt)
            keep.append(xr)
            hargs.append(_ptr(xr))
    ci = ctypes.c_int
    if direction == "fwd":
        args = [_ptr(d_r), _ptr(d_i)] + [ci(k) for k in Kd] + hargs
        args += [ci(j) for j in Jd] + [ci(L), _ptr(tm), ci(M), _ptr(o_r), _ptr(o_i)]
    else:
        args = [_ptr(o_r), _ptr(o_i)] + [ci(k) for k in Kd] + hargs
        args += [ci(j) for j in Jd] + [ci(L), _ptr(tm), ci(M), _ptr(d_r), _ptr(d_i), ci(1)]
    fn(*args)
    out = np.empty(n_out, dtype=cdt)
    out.real = o_r
    out.imag = o_i
    return out


def interp_table(Kd, Jd, L, h, tm, grid, engine="port"):
    """Forward table interpolation of ``grid`` [prod(Kd), reps] -> [M, reps]
    (_nufft.py:998-1097; _interp_table.py:26-159)."""
    f = _interp_port if engine == "port" else _interp_reference
    grid = np.asarray(grid)
    if grid.ndim == 1:
        grid = grid[:, None]
    cols = [f("fwd", Kd, Jd, L, h, tm, grid[:, r]) for r in range(grid.shape[1])]
    return np.stack(cols, axis=1)


def interp_table_adj(Kd, Jd, L, h, tm, samples, engine="port"):
    """Adjoint gridding of ``samples`` [M, reps] -> [prod(Kd), reps]
    (_nufft.py:1101-1191; _interp_table.py:46-191)."""
    f = _interp_port if engine == "port" else _interp_reference
    samples = np.asarray(samples)
    if samples.ndim == 1:
        samples = samples[:, None]
    cols = [f("adj", Kd, Jd, L, h, tm, samples[:, r]) for r in range(samples.shape[1])]
    return np.stack(cols, axis=1)


# ----------------------------------------------------------------------------
# bin sort (new step; CPU restatement is the bit-exact oracle)
# ----------------------------------------------------------------------------
def window_origin(tm, Jd):
    """``koff_d = 1 + floor(tm_d - J_d/2.)`` in double (template.c:865-867)."""
    tm = np.asarray(tm)
    if tm.ndim == 1:
        tm = tm[:, None]
    koff = np.empty(tm.shape, dtype=np.int64)
    for d in range(tm.shape[1]):
        koff[:, d] = 1 + np.floor(tm[:, d].astype(np.float64) - Jd[d] / 2.0)
    return koff


def bin_sort(tm, Jd, Kd, tile):
    """Bin ids, sort keys and stable sort permutation.

    bin id   = tile coordinates of the wrapped window origin, first axis fastest;
    sort key = bin id * prod(tile) + cell index inside the tile (first axis
               fastest); permutation = stable argsort of the keys.
    Returns (bin_ids int32 [M], keys int64 [M], perm int32 [M]).
    """
    koff = window_origin(tm, Jd)
    ndim = koff.shape[1]
    bin_id = np.zeros(koff.shape[0], dtype=np.int64)
    cell = np.zeros(koff.shape[0], dtype=np.int64)
    bstride = 1
    cstride = 1
    for d in range(ndim):
        kw = np.mod(koff[:, d], Kd[d])
        nb = -(-Kd[d] // tile[d])
        bin_id += (kw // tile[d]) * bstride
        cell += (kw % tile[d]) * cstride
        bstride *= nb
        cstride *= tile[d]
    keys = bin_id * cstride + cell
    perm = np.argsort(keys, kind="stable").astype(np.int32)
    return bin_id.astype(np.int32), keys, perm


def forward_slots(keys, perm, tile):
    """Slot list of the paired forward kernel (new preprocessing, no reference counterpart).

    In sorted order, every run of equal sort keys (= samples of one grid cell) is cut
    into pairs (positions 0|1, 2|3, ...; a last odd one stays single); a slot is
    ``(sorted position of its first sample << 1) | has_partner``.  The slots of a bin are
    then ordered by (rank inside the bin's axis-1 column, column), stable, so that
    consecutive slots sit in different columns.  Returns int64 [n_slots].
    """
    ks = np.asarray(keys)[np.asarray(perm)]
    M = ks.shape[0]
    if M == 0:
        return np.zeros(0, dtype=np.int64)
    idx = np.arange(M)
    head = np.ones(M, dtype=bool)
    head[1:] = ks[1:] != ks[:-1]
    runstart = np.maximum.accumulate(np.where(head, idx, 0))
    first = ((idx - runstart) & 1) == 0
    i = idx[first]
    pair = np.zeros(i.shape[0], dtype=np.int64)
    ok = i + 1 < M
    pair[ok] = ks[i[ok] + 1] == ks[i[ok]]
    slots = (i << 1) | pair
    cpt = int(np.prod(tile))
    kk = ks[i]
    bins = kk // cpt
    col = (kk % cpt) % int(tile[0])
    o1 = np.lexsort((col, bins))                    # stable: by bin, then column
    b1, c1 = bins[o1], col[o1]
    g = b1 * int(tile[0]) + c1
    gh = np.ones(g.shape[0], dtype=bool)
    gh[1:] = g[1:] != g[:-1]
    pos = np.arange(g.shape[0])
    rank = pos - np.maximum.accumulate(np.where(gh, pos, 0))
    o2 = np.lexsort((c1, rank, b1))                 # stable: by bin, rank, column
    return slots[o1][o2]


# ----------------------------------------------------------------------------
# operator
# ----------------------------------------------------------------------------
def _as_tuple(seq, typ, n):
    if np.isscalar(seq):
        return (typ(seq),) * n
    if len(seq) != n:
        raise ValueError("array did not have the expected size of {}".format(n))
    return tuple(typ(s) for s in seq)


class OracleNufft(object):
    """Restatement of ``NufftBase`` for the CPU path (_nufft.py:122-935)."""

    def __init__(self, Nd, omega, Jd=4, Kd=None, precision="single",
                 mode="table", Ld=1024, ortho=False, n_shift=None,
                 phasing="real", adjoint_scalefactor=1.0, order="F",
                 engine="port"):
        if np.isscalar(Nd):
            Nd = (Nd,)
        self.Nd = tuple(int(n) for n in Nd)
        self.ndim = len(self.Nd)
        self.Jd = _as_tuple(Jd, int, self.ndim)
        if Kd is None:
            Kd = tuple(int(1.5 * n) for n in self.Nd)
        self.Kd = _as_tuple(Kd, int, self.ndim)
        self.order = order
        self.phasing = phasing
        self.mode = mode
        self.Ld = Ld
        self.ortho = ortho
        self.engine = engine
        self.adjoint_scalefactor = adjoint_scalefactor
        self.scale_ortho = sqrt(int(np.prod(self.Kd))) if ortho else 1
        omega = np.asarray(omega)
        if omega.ndim == 1:
            omega = omega[:, None]
        if omega.shape[1] != self.ndim:
            raise ValueError("number of cols must match NUFFT dimension")
        if precision == "auto":
            precision = "single" if omega.dtype == np.float32 else "double"
        self.precision = precision
        if precision == "single":
            self._real_dtype, self._cplx_dtype = np.dtype(np.float32), np.dtype(np.complex64)
        else:
            self._real_dtype, self._cplx_dtype = np.dtype(np.float64), np.dtype(np.complex128)
        rdt, cdt = self._real_dtype, self._cplx_dtype
        self.n_mid = n_mid_of(self.Nd, phasing)
        if n_shift is None:
            self.n_shift = (0.0,) * self.ndim
        else:
            self.n_shift = _as_tuple(n_shift, float, self.ndim)
        self.M = omega.shape[0]
        # phases are computed from omega BEFORE it is cast (_nufft.py:313-315)
        if phasing == "real":
            # note: at this point the reference's _real_dtype is already set
            self.phase_before = phase_before(self.Kd, self.n_mid, rdt, cdt)
            self.phase_after = phase_after(omega, self.n_mid, self.n_shift, rdt, cdt)
        else:
            self.phase_before = None
            self.phase_after = None
        self.omega = np.asfortranarray(omega.astype(rdt, copy=False))
        self.sn = scaling_factors(self.Nd, self.Kd, self.Jd, phasing).astype(rdt)
        self.phase_shift = None
        if mode == "sparse":
            if phasing == "real":
                # _init_sparsemat calls _set_phase_funcs again (_nufft.py:768-770),
                # now with omega already cast to the precision dtype
                self.phase_before = phase_before(self.Kd, self.n_mid, rdt, cdt)
                self.phase_after = phase_after(self.omega, self.n_mid,
                                               self.n_shift, rdt, cdt)
            self.p = sparse_matrix(self.omega, self.Nd, self.Jd, self.Kd,
                                   phasing, self.n_shift, rdt, cdt)
        elif mode == "table":
            if phasing == "complex" and any(s != 0 for s in self.n_shift):
                self.phase_shift = np.exp(
                    1j * np.dot(self.omega, np.asarray(self.n_shift)))
            self.h = []
            for d in range(self.ndim):
                h = make_table(self.Nd[d], self.Jd[d], self.Kd[d], Ld, phasing)
                if phasing == "complex":
                    h = h.astype(cdt)
                else:
                    h = np.real(h).astype(rdt)
                self.h.append(h)
            tm = np.zeros_like(self.omega)
            for d in range(self.ndim):
                gam = 2 * np.pi / self.Kd[d]
                tm[:, d] = self.omega[:, d] / gam
            self.tm = tm
        else:
            raise ValueError("Invalid NUFFT mode: {}".format(mode))

    # -- helpers mirroring NufftBase._swap_reps/_unswap_reps (:415-423)
    def _swap(self, x, narg):
        if x.size != narg:
            x = x.transpose(tuple(range(1, x.ndim)) + (0,))
        return x

    def _unswap(self, x, narg):
        if x.size != narg:
            x = x.transpose((x.ndim - 1,) + tuple(range(x.ndim - 1)))
        return x

    def interp(self, xk):
        """grid [prod(Kd), reps] -> samples [M, reps] (nufft_forward grid_only)."""
        if self.mode == "table":
            x = interp_table(self.Kd, self.Jd, self.Ld, self.h, self.tm, xk,
                             engine=self.engine)
            if self.phase_shift is not None:
                x = x * self.phase_shift[:, None]
            return x.astype(self._cplx_dtype, copy=False)
        return np.asarray(self.p * xk)

    def interp_adj(self, x):
        """samples [M, reps] -> grid [prod(Kd), reps] (nufft_adj grid_only)."""
        if self.mode == "table":
            if self.phase_shift is not None:
                x = x * self.phase_shift.conj()[:, None]
            x = x.astype(self._cplx_dtype, copy=False)
            return interp_table_adj(self.Kd, self.Jd, self.Ld, self.h, self.tm, x,
                                    engine=self.engine)
        return np.asarray(self.p.conj().T * x)

    def fft(self, x, grid_only=False):
        """_nufft.py:425-452 and :1275-1397."""
        x = np.asarray(x)
        if self.order == "C" and not grid_only:
            x = self._swap(x, int(np.prod(self.Nd)))
        Nd, Kd = self.Nd, self.Kd
        x = np.asfortranarray(x)
        if grid_only:
            x = x.reshape((int(np.prod(Kd)), -1), order="F")
        else:
            x = x.reshape(list(Nd) + [-1], order="F")
        x = x.astype(self._cplx_dtype, copy=False)
        n_reps = x.shape[-1]
        if not grid_only:
            xk = x * self.sn[..., np.newaxis]
            xk = np.fft.fftn(xk, s=Kd, axes=tuple(range(x.ndim - 1)))
            if xk.dtype != self._cplx_dtype:
                xk = xk.astype(self._cplx_dtype)
            if self.phase_before is not None:
                xk *= self.phase_before[..., np.newaxis]
            xk = xk.reshape((int(np.prod(Kd)), n_reps), order="F")
            if self.ortho:
                xk /= self.scale_ortho
        else:
            xk = x
        out = self.interp(xk)
        out = np.reshape(out, (self.M, n_reps), order="F")
        if grid_only:
            return out
        if self.phase_after is not None:
            out = out * self.phase_after[:, None]
        if n_reps == 1:
            out = out[..., 0]
        if self.order == "C":
            out = self._unswap(out, self.M)
        return out

    def adj(self, xk, grid_only=False, return_psf=False):
        """_nufft.py:454-482 and :1459-1578.  ``return_psf`` (:1495,:1517-1518): no
        conj(phase_after), gridded first repetition returned with shape Kd."""
        xk = np.asarray(xk)
        if self.order == "C" and not grid_only:
            xk = self._swap(xk, self.M)
        Nd, Kd = self.Nd, self.Kd
        if xk.size % self.M != 0:
            raise ValueError("invalid size")
        xk = np.asfortranarray(xk).astype(self._cplx_dtype)
        xk = np.reshape(xk, (self.M, -1), order="F").copy(order="A")
        n_reps = xk.shape[-1]
        if self.phase_after is not None and not return_psf:
            xk *= self.phase_after.conj()[:, np.newaxis]
        xk_all = self.interp_adj(xk)
        if grid_only:
            return xk_all
        if xk_all.ndim == 1:
            xk_all = xk_all[:, None]
        xk_all = xk_all.reshape(Kd + (n_reps,), order="F")
        if return_psf:
            return xk_all[..., 0]
        if self.phase_before is not None:
            xk_all = xk_all * self.phase_before.conj()[..., np.newaxis]
        x = np.fft.ifftn(xk_all, s=Kd, axes=tuple(range(xk_all.ndim - 1)))
        if x.dtype != self._cplx_dtype:
            x = x.astype(self._cplx_dtype)
        x = x[tuple([slice(d) for d in Nd] + [slice(None)])]
        if self.ortho:
            x = x * (self.scale_ortho * self.adjoint_scalefactor)
        else:
            x = x * (int(np.prod(Kd)) * self.adjoint_scalefactor)
        x = x * np.conj(self.sn)[..., np.newaxis]
        x = x.astype(self._cplx_dtype, copy=False)
        if n_reps == 1:
            x = x[..., 0]
        if self.order == "C":
            x = self._unswap(x, int(np.prod(Nd)))
        return x

    def norm(self, x):
        """Gram operator adj(fft(x)) (absent upstream)."""
        return self.adj(self.fft(x))


def float64_twin(O):
    """complex128 evaluation of the SAME linear operator as the single-precision oracle ``O``.

    Same float32 coordinates ``tm``, same (float32-accurate) tables, same float32-rounded
    ``sn`` and phases -- only the arithmetic of the interpolation, the FFT and the scalings is
    done in double.  ``rel_l2(y32, twin(x))`` therefore isolates the ROUNDING noise of a
    float32 implementation (the reference's sequential float32 gridding, or the CUDA kernels')
    from everything both implementations share by construction.  Used by the float32 parity
    criterion of tests/golden_util.py:assert_single_parity.  Table mode only.
    """
    if O.precision != "single" or O.mode != "table":
        raise ValueError("float64_twin: single-precision table-mode oracle expected")
    T = OracleNufft(Nd=O.Nd, omega=np.asarray(O.omega, dtype=np.float64), Jd=O.Jd, Kd=O.Kd,
                    precision="double", mode="table", Ld=O.Ld, ortho=O.ortho, n_shift=O.n_shift,
                    phasing=O.phasing, adjoint_scalefactor=O.adjoint_scalefactor, order=O.order,
                    engine=O.engine)
    T.tm = np.asfortranarray(O.tm.astype(np.float64))
    T.h = [np.asarray(h).astype(np.complex128 if np.iscomplexobj(h) else np.float64) for h in O.h]
    T.sn = O.sn.astype(np.float64)
    if O.phase_before is not None:
        T.phase_before = O.phase_before.astype(np.complex128)
    if O.phase_after is not None:
        T.phase_after = O.phase_after.astype(np.complex128)
    if O.phase_shift is not None:
        T.phase_shift = np.asarray(O.phase_shift).astype(np.complex64).astype(np.complex128)
    return T


# ----------------------------------------------------------------------------
# exact transforms for accuracy sanity checks (_dtft.py:16-216)
# ----------------------------------------------------------------------------
def dtft(x, omega, shape, n_shift=None):
    omega = np.asarray(omega, dtype=np.float64)
    dd = omega.shape[1]
    if n_shift is None:
        n_shift = np.zeros(dd)
    x = np.asarray(x).reshape((int(np.prod(shape)), -1), order="F")
    nng = np.meshgrid(*[np.arange(shape[d]) - n_shift[d] for d in range(dd)],
                      indexing="ij")
    ph = np.outer(omega[:, 0], nng[0].ravel(order="F"))
    for d in range(1, dd):
        ph += np.outer(omega[:, d], nng[d].ravel(order="F"))
    out = np.dot(np.exp(-1j * ph), x)
    return out[:, 0] if out.shape[1] == 1 else out


def dtft_adj(xk, omega, shape, n_shift=None):
    omega = np.asarray(omega, dtype=np.float64)
    dd = omega.shape[1]
    if n_shift is None:
        n_shift = np.zeros(dd)
    xk = np.asarray(xk).reshape((omega.shape[0], -1), order="F")
    nng = np.meshgrid(*[np.arange(shape[d]) - n_shift[d] for d in range(dd)],
                      indexing="ij")
    ph = np.outer(nng[0].ravel(order="F"), omega[:, 0])
    for d in range(1, dd):
        ph += np.outer(nng[d].ravel(order="F"), omega[:, d])
    out = np.dot(np.exp(1j * ph), xk)
    out = out.reshape(tuple(shape) + (-1,), order="F")
    return out[..., 0] if out.shape[-1] == 1 else out


# ---------------------------------------------------------------------------------------
# Restatement of the axis-3 FFT pass of mrrt/nufft_b200/csrc/fft_axis3.cuh (option own_fft3):
# the same radix schedule and the same Stockham index arithmetic, in NumPy, so that the
# schedule is pinned against numpy.fft on the CPU (tests/test_oracle.py).  The reference has
# no counterpart (it calls numpy.fft / cuFFT, _nufft.py:1331, :1335-1369).
# ---------------------------------------------------------------------------------------
def axis3_radices(L):
    """Radix schedule of ``axis3_factor``: 8s, 4s, 2s, then 3s, with one (2, 3) pair merged into
    a final radix-6 pass; None if L has another factor."""
    if L < 2:
        return None
    out = []
    while L % 8 == 0:
        out.append(8)
        L //= 8
    while L % 4 == 0:
        out.append(4)
        L //= 4
    n2 = n3 = 0
    while L % 2 == 0:
        n2 += 1
        L //= 2
    while L % 3 == 0:
        n3 += 1
        L //= 3
    if L != 1:
        return None
    six = n2 > 0 and n3 > 0
    if six:
        n2 -= 1
        n3 -= 1
    out += [2] * n2 + [3] * n3 + ([6] if six else [])
    return out


def axis3_stockham(x, inverse=False):
    """Unnormalised DFT of the last axis of ``x`` by the kernel's Stockham passes: butterfly
    ``j`` of a radix-``R`` pass reads rows ``j + r*L/R``, multiplies by
    ``W^(r*k*L/(Ns*R))`` with ``k = j mod Ns`` and writes rows ``(j div Ns)*Ns*R + k + r*Ns``."""
    x = np.asarray(x, dtype=np.complex128)
    L = x.shape[-1]
    rad = axis3_radices(L)
    if rad is None:
        raise ValueError("length %d is not of the form 2^a 3^b" % L)
    sign = 1.0 if inverse else -1.0
    W = np.exp(sign * 2j * np.pi * np.arange(L) / L)
    a = x.copy()
    Ns = 1
    for R in rad:
        b = np.empty_like(a)
        Tn = L // R
        tstep = Tn // Ns
        j = np.arange(Tn)
        k = j % Ns
        j0 = (j // Ns) * Ns * R + k
        v = [a[..., j + r * Tn] * W[(r * k * tstep) % L] for r in range(R)]
        small = np.exp(sign * 2j * np.pi * np.outer(np.arange(R), np.arange(R)) / R)
        for q in range(R):
            b[..., j0 + q * Ns] = sum(small[q, r] * v[r] for r in range(R))
        a = b
        Ns *= R
    return a
