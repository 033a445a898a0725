"""BASELINE configs[4] at its FULL size (3-D 256^3, 3-D radial 102 944 x 512 = 52 707 328 samples,
Kd 384^3, J=6, complex64) through properties that do not need the CPU oracle to run for an
hour.  They tie the full-size operator to the one the oracle CAN check:

  1. the forward transform is sample-wise: its values at every 16th spoke equal the forward
     transform of the operator built on those spokes alone -- the operator that
     tests/test_baseline_configs.py::test_c5_sixteenth_of_the_spokes holds to 1e-5 of the reference;
  2. the adjoint is a sum over samples: with every other sample zeroed it equals the adjoint of
     that same sub-operator;
  3. <A x, A x> = <x, A^H A x> (adjointness without cancellation);
  4. A (x + 2 x') = A x + 2 A x'.

The bookkeeping is run on the CPU first, with the oracle as the operator on a small volume
(-m "not gpu"), so that the full-size GPU run tests the CUDA path and not this file.
(The file sorts after the other GPU tests on purpose: it is the most expensive one.)"""
import numpy as np
import pytest

from golden_util import TOL, rel_l2


def _dot128(a, b, chunk=1 << 22):
    """vdot(a, b) accumulated in complex128 without a full-size float64 copy."""
    a = np.asarray(a).ravel(order="F")          # logical (first-axis-fastest) order for both,
    b = np.asarray(b).ravel(order="F")          # whatever their memory layouts are
    acc = 0j
    for i in range(0, a.size, chunk):
        acc += np.vdot(a[i:i + chunk].astype(np.complex128), b[i:i + chunk].astype(np.complex128))
    return acc


def _image(Nd, seed):
    rs = np.random.RandomState(seed)
    x = rs.standard_normal(Nd).astype(np.float32) + 1j * rs.standard_normal(Nd).astype(np.float32)
    return np.asfortranarray(x.astype(np.complex64))


def subset_and_adjoint_properties(make_op, Nd, om, rows):
    """Errors of the four properties for ``make_op(omega)`` -> operator with fft / adj on NumPy
    arrays; ``rows`` = indices of the samples of the sub-operator."""
    A = make_op(om)
    B = make_op(np.ascontiguousarray(om[rows]))
    x, x2 = _image(Nd, 0), _image(Nd, 1)
    Ax = np.asarray(A.fft(x)).reshape(-1)
    assert Ax.shape == (om.shape[0],) and Ax.dtype == np.complex64
    res = {}
    res["fwd_subset"] = rel_l2(Ax[rows], np.asarray(B.fft(x)).reshape(-1))
    y = np.zeros_like(Ax)
    y[rows] = Ax[rows]
    res["adj_subset"] = rel_l2(np.asarray(A.adj(y)), np.asarray(B.adj(np.ascontiguousarray(Ax[rows]))))
    AhAx = np.asarray(A.adj(Ax))
    assert AhAx.shape == tuple(Nd)
    lhs = _dot128(Ax, Ax)
    rhs = _dot128(x, AhAx)
    res["adjointness"] = abs(lhs - rhs) / abs(lhs)
    Ax2 = np.asarray(A.fft(x2)).reshape(-1)
    res["linearity"] = rel_l2(np.asarray(A.fft(np.asfortranarray(x + 2 * x2))).reshape(-1), Ax + 2 * Ax2)
    return res


def _every_nth_spoke_rows(spokes, nread, nth):
    idx = np.arange(0, spokes, nth)
    return (idx[:, None] * nread + np.arange(nread)[None, :]).ravel()


def test_property_bookkeeping_with_the_oracle():
    """The helper itself, on the CPU: 3-D radial 96 x 32 on a 16^3 volume with the oracle as the
    operator.  Forward values of a sample subset are identical, the masked adjoint agrees to
    float32 summation order, adjointness and linearity hold to float32 rounding."""
    import bench
    from oracle import nufft_oracle as orc

    Nd, Kd = (16, 16, 16), (24, 24, 24)
    om = bench.radial3d(96, 32)
    rows = _every_nth_spoke_rows(96, 32, 4)
    assert rows.size == 24 * 32 and np.array_equal(om[rows][:32], om[:32])

    def make(o):
        return orc.OracleNufft(Nd=Nd, omega=o, Jd=6, Kd=Kd, precision="single", mode="table")

    res = subset_and_adjoint_properties(make, Nd, om, rows)
    assert res["fwd_subset"] <= 1e-7, res
    assert res["adj_subset"] <= 2e-6, res
    assert res["adjointness"] <= 2e-6, res
    assert res["linearity"] <= 2e-6, res


@pytest.mark.gpu
def test_bench_workload_full_size_properties():
    """The four properties on the bench operator itself (52.7 M samples, the column-group
    adjoint and the tiled forward with the own FFT passes), at north_star's complex64 tolerance."""
    import bench
    from mrrt.nufft_b200 import NufftBase

    om = bench.radial3d(bench.SPOKES, bench.NREAD)
    rows = _every_nth_spoke_rows(bench.SPOKES, bench.NREAD, 16)
    ops = []

    def make(o):
        ops.append(NufftBase(Nd=bench.ND, omega=o, Jd=bench.JD, Kd=bench.KD, precision="single",
                             mode="table"))
        return ops[-1]

    res = subset_and_adjoint_properties(make, bench.ND, om, rows)
    full = ops[0]
    assert full.M == bench.SPOKES * bench.NREAD
    assert full.option("last_fwd_kernel") == 1 and full.option("last_adj_kernel") == 5
    print("full-size C5: " + ", ".join("%s %.2e" % kv for kv in sorted(res.items())))
    for name, err in res.items():
        assert err <= TOL["single"], res
