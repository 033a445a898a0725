"""Generate the golden vectors in tests/golden/ by running the REAL reference
package (mrrt.nufft.NufftBase, CPU path) in this container.

    oracle/build_reference_scratch.sh /tmp/refbuild
    PYTHONPATH=/tmp/refbuild python tests/golden/make_golden.py

The scratch build lives outside the repo; only the small .npz fixtures and this
script are committed.  The cases are the reference's own test configurations
(tests/test_nufft.py:60-324: 1-D N=64/K=128/J=6, 2-D 16x16 with even+odd K and J,
3-D 8^3, the odd-shift adjoint case) over {table, sparse} x {single, double} x
{real, complex} phasing, plus a few mid-size cases shaped like BASELINE.json's
configs.  Each case stores inputs, plan arrays (tm, sn, koff windows), the full
fft/adj outputs and the grid_only interpolation outputs.  Lookup tables are stored
once per distinct (N, K, J, L, phasing) in tables.npz (float32/complex64 is lossless:
the reference's tables are float32-accurate in every precision).
"""
import json
import os
import sys
import warnings

import numpy as np

warnings.simplefilter("ignore")

from mrrt.nufft import NufftBase  # noqa: E402  (the reference)
from mrrt.nufft._nufft import nufft_forward, nufft_adj  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def uniform_freqs(Nd):
    fs = [2 * np.pi * np.arange(n) / n for n in Nd]
    fs = np.meshgrid(*fs, indexing="ij")
    return np.hstack([f.reshape((-1, 1), order="F") for f in fs])


def perturbed_gridpoints(Nd, rel_std=0.5, seed=1234):
    """tests/test_nufft.py:20-44 (the reference's trajectory fixture)."""
    rstate = np.random.RandomState(seed)
    Nd = np.atleast_1d(np.asarray(Nd))
    omega = uniform_freqs(Nd)
    df = 2 * np.pi / Nd
    npoints = omega.shape[0]
    for d in range(len(Nd)):
        omega[:, d] += df[d] * rel_std * rstate.rand(npoints)
        omega[:, d] += np.min(omega[:, d])
        omega[:, d] *= (2 * np.pi) / np.max(omega[:, d])
    return omega


def radial2d(S, n):
    ang = np.pi * np.arange(S) / S
    r = 2 * np.pi * (np.arange(n) - n / 2) / n
    return np.stack([np.outer(np.cos(ang), r).ravel(),
                     np.outer(np.sin(ang), r).ravel()], axis=1)


def grid_only_inputs(seed, PK, M, n_reps, cdt):
    """Inputs of the grid_only checks; regenerated (not stored) by the tests."""
    rs = np.random.RandomState(seed + 1)
    g = rs.standard_normal((PK, n_reps)) + 1j * rs.standard_normal((PK, n_reps))
    ysamp = rs.standard_normal((M, n_reps)) + 1j * rs.standard_normal((M, n_reps))
    return g.astype(cdt), ysamp.astype(cdt)


def run_case(name, Nd, Kd, Jd, omega, n_shift, mode, precision, phasing,
             order="F", Ld=1024, n_reps=1, ortho=False, adjoint_scalefactor=1.0,
             tables=None, seed=1234, store_p=False):
    A = NufftBase(omega=omega, Nd=Nd, Jd=Jd, Kd=Kd, n_shift=n_shift, mode=mode,
                  Ld=Ld, precision=precision, phasing=phasing, order=order,
                  ortho=ortho, adjoint_scalefactor=adjoint_scalefactor)
    rs = np.random.RandomState(seed)
    Nd_t = tuple(A.Nd)
    shape = Nd_t + (n_reps,) if n_reps > 1 else Nd_t
    x = rs.standard_normal(shape) + 1j * rs.standard_normal(shape)
    if order == "C" and n_reps > 1:
        x = np.moveaxis(x, -1, 0)
    y = A.fft(x)
    x_adj = A.adj(y)
    # interpolation only (grid_only switches, _nufft.py:1302-1304,1507-1509)
    PK = int(np.prod(A.Kd))
    g, ysamp = grid_only_inputs(seed, PK, A.M, n_reps, A._cplx_dtype)
    interp_out = nufft_forward(A, g.copy(), grid_only=True)
    grid_out = nufft_adj(A, ysamp.copy(), grid_only=True)
    grid_out = np.asarray(grid_out).reshape((PK, n_reps), order="F")
    cfg = dict(name=name, Nd=list(map(int, A.Nd)), Kd=list(map(int, A.Kd)),
               Jd=list(map(int, A.Jd)), Ld=int(Ld), mode=mode, precision=precision,
               phasing=phasing, order=order, n_reps=n_reps, ortho=bool(ortho),
               adjoint_scalefactor=float(adjoint_scalefactor),
               n_shift=[float(s) for s in A.n_shift], seed=seed)
    out = dict(cfg=json.dumps(cfg), omega=np.asarray(omega, dtype=np.float64)
               if np.asarray(omega).dtype != np.float32 else np.asarray(omega),
               x=x, y=y, x_adj=x_adj, interp_out=interp_out,
               grid_out=grid_out, sn=A.sn)
    if A.phase_after is not None:
        out["phase_after"] = A.phase_after
    if mode == "table":
        out["tm"] = A.tm
        for d in range(A.ndim):
            key = "N%d_K%d_J%d_L%d_%s" % (A.Nd[d], A.Kd[d], A.Jd[d], Ld, phasing)
            h = np.asarray(A.h[d])
            h32 = h.astype(np.complex64 if np.iscomplexobj(h) else np.float32)
            assert np.array_equal(h32.astype(h.dtype), h), "table not f32-exact"
            if key in tables:
                assert np.array_equal(tables[key], h32)
            tables[key] = h32
    elif store_p:
        p = A.p.tocsr()
        p.sort_indices()
        out["p_indptr"] = p.indptr.astype(np.int32)
        out["p_indices"] = p.indices.astype(np.int32)
        out["p_data"] = p.data
    return out


def main():
    tables = {}
    cases = {}
    modes = ["table", "sparse"]
    precs = ["single", "double"]
    phs = ["real", "complex"]

    # 1-D (tests/test_nufft.py:99-169)
    om1 = perturbed_gridpoints(64)
    for mode in modes:
        for prec in precs:
            for ph in phs:
                nm = "d1_%s_%s_%s" % (mode, prec, ph)
                cases[nm] = run_case(nm, 64, 128, 6, om1, 32, mode, prec, ph,
                                     tables=tables, store_p=True)
    nm = "d1_table_single_real_C2"
    cases[nm] = run_case(nm, 64, 128, 6, om1, 32, "table", "single", "real",
                         order="C", n_reps=2, tables=tables)

    # 2-D (tests/test_nufft.py:172-244)
    om2 = perturbed_gridpoints((16, 16))
    for mode in modes:
        for prec in precs:
            for ph in phs:
                for Kd in [(32, 32), (33, 31)]:
                    for Jd in [6, 7]:
                        nm = "d2_%s_%s_%s_K%d_J%d" % (mode, prec, ph, Kd[0], Jd)
                        cases[nm] = run_case(nm, (16, 16), Kd, Jd, om2, (8.0, 8.0),
                                             mode, prec, ph, tables=tables)
    nm = "d2_table_double_real_F2"
    cases[nm] = run_case(nm, (16, 16), (32, 32), 6, om2, (8.0, 8.0), "table",
                         "double", "real", n_reps=2, tables=tables)
    nm = "d2_table_single_real_C2"
    cases[nm] = run_case(nm, (16, 16), (32, 32), 6, om2, (8.0, 8.0), "table",
                         "single", "real", order="C", n_reps=2, tables=tables)

    # 3-D (tests/test_nufft.py:247-324)
    om3 = perturbed_gridpoints((8, 8, 8))
    for mode in modes:
        for prec in precs:
            for ph in phs:
                nm = "d3_%s_%s_%s" % (mode, prec, ph)
                cases[nm] = run_case(nm, (8, 8, 8), (16, 16, 16), 6, om3,
                                     (4.0, 4.0, 4.0), mode, prec, ph, n_reps=4
                                     if (mode, prec, ph) == ("table", "single", "real")
                                     else 1, tables=tables)

    # odd-shift adjoint case (tests/test_nufft.py:60-96)
    o1 = 2 * np.pi * np.array([0.0, 0.1, 0.3, 0.4, 0.7, 0.9])
    om_adj = np.stack((o1, o1[::-1].copy()), axis=-1)
    for mode in modes:
        for ph in phs:
            nm = "adjshift_%s_%s" % (mode, ph)
            cases[nm] = run_case(nm, (4, 8), (8, 16), (8, 8), om_adj, [2.7, 3.1],
                                 mode, "single", ph, n_reps=3, tables=tables,
                                 store_p=True)

    # ortho / adjoint_scalefactor / default Kd=1.5N / unequal J
    nm = "d2_table_double_real_ortho"
    cases[nm] = run_case(nm, (16, 16), (32, 32), 6, om2, (8.0, 8.0), "table",
                         "double", "real", ortho=True, adjoint_scalefactor=0.5,
                         tables=tables)
    rs = np.random.RandomState(7)
    om_r3 = (rs.rand(2000, 3) * 2 - 1) * np.pi
    for prec in precs:
        nm = "d3_mid_table_%s_real_J4" % prec
        cases[nm] = run_case(nm, (16, 16, 16), None, 4, om_r3, None, "table", prec,
                             "real", tables=tables)
    nm = "d3_mid_table_single_real_J546"
    cases[nm] = run_case(nm, (12, 16, 10), (24, 25, 18), (5, 4, 6), om_r3, None,
                         "table", "single", "real", tables=tables)

    # mid-size 2-D radial, shaped like BASELINE configs[0] (scaled down) -- omega
    # handed over in float32 as the bench does
    om_rad = radial2d(40, 96)
    for prec in precs:
        nm = "d2_radial_table_%s_real" % prec
        om = om_rad.astype(np.float32 if prec == "single" else np.float64)
        cases[nm] = run_case(nm, (48, 48), (96, 96), 6, om, None, "table", prec,
                             "real", tables=tables)
    rs = np.random.RandomState(1)
    om_rd = np.clip((np.pi / 3) * rs.standard_normal((3000, 2)), -np.pi, np.pi - 1e-6)
    for mode in modes:
        nm = "d2_randdens_%s_double_real" % mode
        cases[nm] = run_case(nm, (48, 48), (96, 96), 6, om_rd, None, mode,
                             "double", "real", tables=tables)

    np.savez_compressed(os.path.join(HERE, "tables.npz"), **tables)
    for nm, c in cases.items():
        np.savez_compressed(os.path.join(HERE, nm + ".npz"), **c)
    tot = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE)
              if f.endswith(".npz"))
    print("wrote %d cases, %d tables, %.2f MB" % (len(cases), len(tables), tot / 1e6))


if __name__ == "__main__":
    sys.exit(main())
