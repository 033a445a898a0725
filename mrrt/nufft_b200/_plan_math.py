"""Host-side (NumPy) plan constants: lookup tables, deapodization factors, phases and the
per-axis sparse-mode coefficients.  Values are formed with the same floating-point
operations, in the same order and dtypes, as the reference so that they are
bit-identical to it (checked against tests/golden/ which the real reference produced).
"""
import numpy as np

from ._kernels import BeattyKernel, kaiser_bessel, kaiser_bessel_ft

TWO_PI = 2 * np.pi


def real_cplx_dtypes(precision):
    if precision == "single":
        return np.dtype(np.float32), np.dtype(np.complex64)
    if precision == "double":
        return np.dtype(np.float64), np.dtype(np.complex128)
    raise ValueError("precision must be 'single', 'double' or 'auto'")


def n_mid(Nd, phasing):
    """Midpoint of the scaling factors (_nufft.py:623-628)."""
    if phasing == "real":
        return tuple(n // 2 for n in Nd)
    return tuple((n - 1) / 2.0 for n in Nd)


def axis_window(om, J, K):
    """Per-sample window origin and tap arguments for one axis.

    ``om`` is a 1-D array already in the precision dtype.  Returns ``koff`` (float array,
    ``floor(om/gam - J/2)``, _utils.py:20-43) and ``arg`` [J, M] (``-j + (om/gam - koff)``,
    j = 1..J, promoted to float64 by the integer tap index, _utils.py:72-78).
    """
    gam = TWO_PI / K
    tt = om / gam
    koff = np.floor(tt - J / 2.0)
    dk = tt - koff
    arg = -np.arange(1, J + 1)[:, None] + dk[None, :]
    return koff, arg


def axis_coefficients(om, N, J, K, alpha, phasing):
    """Coefficients u [J, M] (float64 / complex128) and wrapped columns kd [J, M] of one
    axis of the interpolation matrix (_nufft.py:772-802)."""
    koff, arg = axis_window(om, J, K)
    u = kaiser_bessel(arg, J=J, alpha=alpha, m=0)
    if phasing == "complex":
        u = np.exp((1j * (TWO_PI / K) * (N - 1) / 2.0) * arg) * u
    kd = np.mod(np.arange(1, J + 1)[:, None] + koff[None, :], K)
    return u, kd


def lookup_table(N, J, K, L, phasing):
    """Length ``J*L+1`` lookup table of one axis.

    The reference obtains it from a dummy 1-D *single precision* sparse operator sampled
    at ``L`` offsets and reads matrix columns ``J-1 .. 0`` (_nufft.py:1195-1243,
    how="fast"); the values are therefore float32-accurate in every precision.  Here the
    same J columns are assembled directly.
    """
    steps = np.arange(L) if N % 2 == 0 else np.arange(1, L + 1)
    t1 = J / 2.0 - 1 + steps / L
    om = (t1 * 2 * np.pi / K).astype(np.float32)
    alpha = BeattyKernel.beatty_alpha(J, K, N)
    u, kd = axis_coefficients(om, N, J, K, alpha, phasing)
    if phasing == "complex":
        u = u.conj().astype(np.complex64)
    else:
        u = u.astype(np.float32)
    cols = np.zeros((L, J), dtype=u.dtype)
    rows = np.broadcast_to(np.arange(L)[None, :], kd.shape)
    keep = kd < J
    np.add.at(cols, (rows[keep], kd[keep].astype(np.intp)), u[keep])
    h = cols[:, ::-1].ravel(order="F")
    if N % 2 == 0:
        return np.concatenate((h, h[:1]))
    return np.concatenate((h[-1:], h))


def deapodization_1d(Nd, Kd, Jd, alphas, phasing):
    """Per-axis image-domain scaling ``1/KBFT((n - n_mid)/K)`` (_nufft.py:737-746)."""
    mids = n_mid(Nd, phasing)
    out = []
    for N, K, J, a, mid in zip(Nd, Kd, Jd, alphas, mids):
        nc = np.arange(-mid, -mid + N)
        out.append(1 / kaiser_bessel_ft(nc / K, J, a, 0, 1))
    return out


def dense_sn(sn1d, Nd):
    """The reference's dense ``sn`` (outer products in float64, _nufft.py:747-748)."""
    sn = np.array([1.0])
    for f in sn1d:
        sn = np.outer(sn.ravel(), f.conj())
    return sn.reshape(Nd)


def phase_before_angles(Kd, mids, rdt):
    """Per-axis angle ``(2 pi/K n_mid) k`` in the precision dtype (_nufft.py:708-710)."""
    return [(2 * np.pi / K * mid) * np.arange(K, dtype=rdt) for K, mid in zip(Kd, mids)]


def dense_phase_before(angles, cdt):
    """_nufft.py:711-715: angles outer-summed in the precision dtype, then exp."""
    phase = angles[0]
    for d in range(1, len(angles)):
        phase = phase.reshape(phase.shape + (1,)) + angles[d].reshape((1,) * d + (angles[d].size,))
    return np.exp(1j * phase).astype(cdt, copy=False)


def _phase_after_block(omega, shift, cdt):
    return np.exp(1j * np.dot(omega, shift)).astype(cdt, copy=False)


def phase_after(omega, mids, n_shift, rdt, cdt):
    """``exp(i omega.(n_shift - n_mid))`` (_nufft.py:717-724).

    Stays on the HOST on purpose: in single precision the angle is a float32 dot product whose
    rounding (~3e-5 rad at |angle| ~ 500) is part of what the reference computes, and only the
    same NumPy expression on the same host reproduces it bit for bit (a device evaluation would
    differ by an ulp of the angle, i.e. by more than the 1e-5 parity budget).  Large sample sets
    are evaluated in row blocks on a thread pool (NumPy releases the GIL inside the loops; every
    row's result is independent of the blocking)."""
    shift_vec = [(s - m) for s, m in zip(n_shift, mids)]
    shift = np.asarray(shift_vec, dtype=rdt)
    M = omega.shape[0]
    nblk = min(32, M // (1 << 18))
    if nblk < 2:
        return _phase_after_block(omega, shift, cdt)
    import os
    from concurrent.futures import ThreadPoolExecutor

    out = np.empty(M, dtype=cdt)
    edges = np.linspace(0, M, nblk + 1).astype(np.int64)

    def run(b):
        out[edges[b]:edges[b + 1]] = _phase_after_block(omega[edges[b]:edges[b + 1]], shift, cdt)

    nthr = min(nblk, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else 4)
    with ThreadPoolExecutor(max_workers=max(1, nthr)) as ex:
        list(ex.map(run, range(nblk)))
    return out
