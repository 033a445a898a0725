"""Pin the CPU oracle (oracle/) against the golden vectors produced by the real
reference package, and the plain-C restatement against the reference's own compiled C.
CPU only."""
import numpy as np
import pytest

from golden_util import (PSF_CASES, TOL, case_names, ctor_kwargs, grid_only_inputs, load_case,
                         psf_cases, rel_l2, table_key, tables)
from oracle import nufft_oracle as orc

ENGINES = ["port"] + (["reference"] if orc.have_reference_engine() else [])
# the oracle is the same algorithm on the same NumPy: only -ffast-math reassociation
# in the reference build separates them
ORACLE_TOL = {"single": 2e-6, "double": 1e-13}


@pytest.mark.parametrize("name", case_names())
def test_oracle_matches_golden(name):
    cfg, z = load_case(name)
    tol = ORACLE_TOL[cfg["precision"]]
    for engine in ENGINES:
        A = orc.OracleNufft(omega=z["omega"], engine=engine, **ctor_kwargs(cfg))
        assert np.array_equal(A.sn, z["sn"])
        if "phase_after" in z:
            assert np.array_equal(A.phase_after, z["phase_after"])
        if cfg["mode"] == "table":
            assert np.array_equal(A.tm, z["tm"])
            for d in range(A.ndim):
                key = table_key(A.Nd[d], A.Kd[d], A.Jd[d], cfg["Ld"], cfg["phasing"])
                assert np.array_equal(A.h[d], tables()[key].astype(A.h[d].dtype))
        elif "p_data" in z:
            p = A.p.tocsr()
            p.sort_indices()
            assert np.array_equal(p.indptr, z["p_indptr"])
            assert np.array_equal(p.indices, z["p_indices"])
            assert np.array_equal(p.data, z["p_data"])
        y = A.fft(z["x"])
        assert y.dtype == z["y"].dtype
        assert rel_l2(y, z["y"]) <= tol
        xa = A.adj(z["y"])
        assert rel_l2(xa, z["x_adj"]) <= tol
        g, ysamp = grid_only_inputs(cfg["seed"], int(np.prod(A.Kd)), A.M,
                                    cfg["n_reps"], A._cplx_dtype)
        assert rel_l2(A.fft(g, grid_only=True), z["interp_out"]) <= tol
        assert rel_l2(A.adj(ysamp, grid_only=True), z["grid_out"]) <= tol


@pytest.mark.parametrize("name", PSF_CASES)
def test_oracle_return_psf_matches_golden(name):
    """adj(..., return_psf=True): no conj(phase_after), first repetition, shape Kd."""
    cfg, z = load_case(name)
    want = psf_cases()[name]
    A = orc.OracleNufft(omega=z["omega"], **ctor_kwargs(cfg))
    _, ysamp = grid_only_inputs(cfg["seed"], int(np.prod(A.Kd)), A.M, cfg["n_reps"],
                                A._cplx_dtype)
    got = A.adj(ysamp, return_psf=True)
    assert got.shape == want.shape and got.dtype == want.dtype
    assert rel_l2(got, want) <= ORACLE_TOL[cfg["precision"]]


@pytest.mark.skipif(not orc.have_reference_engine(), reason="oracle/_ref not built")
@pytest.mark.parametrize("ndim,cplx,rdt", [(n, c, r) for n in (1, 2, 3)
                                          for c in (False, True)
                                          for r in (np.float32, np.float64)])
def test_port_matches_reference_c(ndim, cplx, rdt):
    """Plain-C restatement vs the reference translation unit compiled as-is."""
    rs = np.random.RandomState(ndim)
    Kd = (20, 18, 15)[:ndim]
    Jd = (6, 5, 4)[:ndim]
    L = 64
    M = 500
    cdt = np.complex64 if rdt == np.float32 else np.complex128
    h = []
    for J in Jd:
        t = rs.standard_normal(J * L + 1)
        if cplx:
            t = t + 1j * rs.standard_normal(J * L + 1)
        h.append(t.astype(cdt if cplx else rdt))
    tm = np.asfortranarray((rs.rand(M, ndim) * 3 - 1) * np.asarray(Kd)).astype(rdt)
    g = (rs.standard_normal(int(np.prod(Kd))) + 1j * rs.standard_normal(int(np.prod(Kd)))).astype(cdt)
    s = (rs.standard_normal(M) + 1j * rs.standard_normal(M)).astype(cdt)
    tol = 5e-6 if rdt == np.float32 else 1e-13
    a = orc.interp_table(Kd, Jd, L, h, tm, g, engine="port")
    b = orc.interp_table(Kd, Jd, L, h, tm, g, engine="reference")
    assert rel_l2(a, b) <= tol
    a = orc.interp_table_adj(Kd, Jd, L, h, tm, s, engine="port")
    b = orc.interp_table_adj(Kd, Jd, L, h, tm, s, engine="reference")
    assert rel_l2(a, b) <= tol


def test_nufft_offset_known_answers():
    """The reference's own known-answer test (tests/test_utils.py:15-22)."""
    assert orc.nufft_offset(0, 4, 128) == -2
    assert orc.nufft_offset(0, 5, 128) == -3
    assert orc.nufft_offset(0, 5.5, 128) == -3
    assert orc.nufft_offset(0, 6, 128) == -3
    assert orc.nufft_offset(np.asarray([0]), 6.01, 128) == np.asarray([-4])


def test_oracle_vs_exact_dtft():
    """NUFFT accuracy against the exact transform (tests/test_nufft.py:172-244)."""
    cfg, z = load_case("d2_table_double_real_K32_J6")
    A = orc.OracleNufft(omega=z["omega"], **ctor_kwargs(cfg))
    y_true = orc.dtft(z["x"], z["omega"], A.Nd, A.n_shift)
    np.testing.assert_allclose(A.fft(z["x"]), y_true, rtol=1e-3, atol=1e-5)
    x_true = orc.dtft_adj(z["y"], z["omega"], A.Nd, A.n_shift)
    np.testing.assert_allclose(A.adj(z["y"]), x_true, rtol=1e-3, atol=1e-5)


def test_bin_sort_properties():
    rs = np.random.RandomState(3)
    Kd, Jd, tile = (40, 36, 30), (6, 6, 4), (8, 8, 4)
    tm = ((rs.rand(5000, 3) * 2 - 1) * np.asarray(Kd)).astype(np.float32)
    bins, keys, perm = orc.bin_sort(tm, Jd, Kd, tile)
    assert np.array_equal(np.sort(perm), np.arange(5000))
    assert np.all(np.diff(keys[perm]) >= 0)
    # stability: equal keys keep acquisition order
    same = np.diff(keys[perm]) == 0
    assert np.all(np.diff(perm)[same] > 0)
    nb = [-(-k // t) for k, t in zip(Kd, tile)]
    assert bins.min() >= 0 and bins.max() < np.prod(nb)


@pytest.mark.parametrize("L", [384, 192, 96, 54, 48, 36, 12, 8, 6, 2])
def test_axis3_stockham_schedule(L):
    """The radix schedule / index arithmetic of the own axis-3 FFT pass (csrc/fft_axis3.cuh,
    restated in the oracle) is a DFT: against numpy.fft in both directions."""
    rs = np.random.RandomState(L)
    x = rs.standard_normal((3, L)) + 1j * rs.standard_normal((3, L))
    assert np.allclose(orc.axis3_stockham(x), np.fft.fft(x, axis=-1), rtol=0, atol=1e-11)
    assert np.allclose(orc.axis3_stockham(x, inverse=True), np.fft.ifft(x, axis=-1) * L, rtol=0, atol=1e-11)
    assert orc.axis3_radices(35) is None and orc.axis3_radices(384) == [8, 8, 6]
