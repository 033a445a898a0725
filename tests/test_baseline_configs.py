"""Parity at the BASELINE.json configs' OWN sizes (C1..C5), CUDA path
through the C ABI against the reference's compiled C interpolators (``oracle/_ref``, driven by
the oracle's restatement of the reference pipeline; the restated C when ``_ref`` is absent).

  C1  2-D 256^2, radial 402 x 512, Kd 512^2, J=6, table, complex64, 1 coil      (full size)
  C2  2-D 256^2, random density M=205 824, Kd 512^2, J=6, complex128, table + sparse (full)
  C3  3-D 128^3 stack-of-stars 201 x 256 x 128 (M = 6 586 368), Kd 192^3, J=4, c64  (full)
  C4  2-D 320^2, radial 503 x 640, Kd 480^2, J=6, 32 coils, complex64              (full)
  C5  3-D 256^3, 3-D radial 102 944 x 512, Kd 384^3, J=6, c64: every 16th spoke
      (M = 3 294 208); the full trajectory is checked through
      size-independent properties in test_gpu_parity.py::test_adjointness_and_linearity_large
      and tests/test_workload_full_size.py.

Tolerances are north_star's: rel-L2 <= 1e-5 (complex64), <= 1e-12 (complex128); the float32
ADJOINT uses golden_util.assert_single_parity (1e-5 against the reference, or -- where two
float32 accumulations cannot agree to 1e-5 -- 1e-5 against the float64 evaluation of the same
operator and no worse than the reference itself)."""
import numpy as np
import pytest

from golden_util import TOL, assert_single_parity, rel_l2

pytestmark = pytest.mark.gpu


def _engine():
    from oracle import nufft_oracle as orc

    return "reference" if orc.have_reference_engine() else "port"


def radial2d(S, n, dtype):
    """ang_s = pi*s/S, r_i = 2*pi*(i - n/2)/n."""
    ang = np.pi * np.arange(S) / S
    r = 2 * np.pi * (np.arange(n) - n / 2) / n
    return np.stack([np.outer(np.cos(ang), r).ravel(), np.outer(np.sin(ang), r).ravel()], 1).astype(dtype)


def stack_of_stars(S, n, P, dtype):
    """2-D radial in (omega1, omega2) x omega3 = 2*pi*(p - P/2)/P (on the grid of axis 3)."""
    r2 = radial2d(S, n, np.float64)
    kz = 2 * np.pi * (np.arange(P) - P // 2) / P
    return np.concatenate([np.concatenate([r2, np.full((r2.shape[0], 1), z)], 1) for z in kz],
                          0).astype(dtype)


def _image(shape, seed, cdt):
    rs = np.random.RandomState(seed)
    return (rs.standard_normal(shape) + 1j * rs.standard_normal(shape)).astype(cdt)


def _single_case(Nd, Kd, J, om, ncoil=1, fwd_kernel=None, adj_kernel=None):
    """fft and adj of a complex64 table-mode operator against the reference + float64 twin."""
    from oracle import nufft_oracle as orc
    from mrrt.nufft_b200 import NufftBase

    A = NufftBase(Nd=Nd, omega=om, Jd=J, Kd=Kd, precision="single", mode="table")
    O = orc.OracleNufft(Nd=Nd, omega=om, Jd=J, Kd=Kd, precision="single", mode="table",
                        engine=_engine())
    assert np.array_equal(A.tm.cpu().numpy(), O.tm)
    shape = tuple(Nd) + ((ncoil,) if ncoil > 1 else ())
    x = _image(shape, 0, np.complex64)
    yo = O.fft(x)
    y = A.fft(x)
    e_f = rel_l2(y, yo)
    assert e_f <= TOL["single"], e_f
    if fwd_kernel is not None:
        assert A.option("last_fwd_kernel") == fwd_kernel
    xo = O.adj(yo)
    xa = A.adj(yo)
    if adj_kernel is not None:
        assert A.option("last_adj_kernel") == adj_kernel
    x64 = orc.float64_twin(O).adj(yo.astype(np.complex128))
    return (e_f,) + assert_single_parity(xa, xo, x64, "adjoint")


def test_c1_full_2d_radial_single_coil():
    """BASELINE configs[0] at its own size; one coil, so the single-coil 2-D adjoint kernel
    is the one that runs (reference hot loop: template.c:417-531)."""
    res = _single_case((256, 256), (512, 512), 6, radial2d(402, 512, np.float32),
                       fwd_kernel=1, adj_kernel=4)
    print("C1 fft %.2e; adj vs ref %.2e, vs f64: cuda %.2e ref %.2e" % res)


@pytest.mark.parametrize("mode", ["table", "sparse"])
def test_c2_full_2d_random_density_double(mode):
    """BASELINE configs[1] at its own size: complex128, table mode against the reference's
    table path and sparse mode against its sparse path (trap 2)."""
    from oracle import nufft_oracle as orc
    from mrrt.nufft_b200 import NufftBase

    rs = np.random.RandomState(1)
    om = np.clip((np.pi / 3) * rs.standard_normal((205824, 2)), -np.pi, np.pi - 1e-6)
    Nd, Kd = (256, 256), (512, 512)
    A = NufftBase(Nd=Nd, omega=om, Jd=6, Kd=Kd, precision="double", mode=mode)
    O = orc.OracleNufft(Nd=Nd, omega=om, Jd=6, Kd=Kd, precision="double", mode=mode,
                        engine=_engine())
    x = _image(Nd, 2, np.complex128)
    yo = O.fft(x)
    assert rel_l2(A.fft(x), yo) <= TOL["double"]
    assert rel_l2(A.adj(yo), O.adj(yo)) <= TOL["double"]
    if mode == "table":
        assert A.option("last_fwd_kernel") == 1 and A.option("last_adj_kernel") == 4


def test_c3_full_3d_stack_of_stars():
    """BASELINE configs[2] at its own size.  omega3 sits ON the grid of axis 3 (t - k integer,
    the alf == 0 / h[J*L] corner of template.c:870-873) for every sample."""
    res = _single_case((128, 128, 128), (192, 192, 192), 4, stack_of_stars(201, 256, 128, np.float32),
                       fwd_kernel=1, adj_kernel=5)
    print("C3 fft %.2e; adj vs ref %.2e, vs f64: cuda %.2e ref %.2e" % res)


def test_c4_full_2d_32_coils():
    """BASELINE configs[3] at its own size: one full coil set (32 coils)."""
    res = _single_case((320, 320), (480, 480), 6, radial2d(503, 640, np.float32), ncoil=32,
                       fwd_kernel=1, adj_kernel=4)
    print("C4 fft %.2e; adj vs ref %.2e, vs f64: cuda %.2e ref %.2e" % res)


def test_c5_sixteenth_of_the_spokes():
    """BASELINE configs[4] on every 16th spoke of the bench trajectory."""
    import bench

    idx = np.arange(0, bench.SPOKES, 16)
    om = np.concatenate([bench.radial3d(bench.SPOKES, bench.NREAD, int(s), int(s) + 1) for s in idx], 0)
    res = _single_case(bench.ND, bench.KD, bench.JD, om, fwd_kernel=1, adj_kernel=5)
    print("C5/16 fft %.2e; adj vs ref %.2e, vs f64: cuda %.2e ref %.2e" % res)
