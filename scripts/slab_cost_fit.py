"""Fit of the slab cost model of SlabShardedNufft on ONE GPU: the slabs of a G-rank run are built
one after the other, their per-slab stages (axis-3 pass forward + adjoint, interpolation forward +
adjoint) are timed with CUDA events, and  t = a * samples + b * occupied cells + c * rows  is
fitted by least squares.  CELL_COST = b / a, row cost = c / a (in "samples").  Writes
profiles/r02_slab_cost_fit.md.   usage: python scripts/slab_cost_fit.py [G ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mrrt.nufft_b200 import _slab  # noqa: E402
from mrrt.nufft_b200._slab import CudaSlabKernels, row_statistics, slab_boundaries  # noqa: E402


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    Gs = [int(a) for a in sys.argv[1:]] or [8, 4]
    Nd, Kd, J = bench.ND, bench.KD, bench.JD
    K1, K2, K3 = Kd
    om = bench.radial3d(bench.SPOKES, bench.NREAD)
    rdt = np.dtype(np.float32)
    rows, n_row, cells_row = row_statistics(om, (J, J, J), Kd, rdt, torch.device("cuda"))
    row_cost = _slab.ROW_COST_PER_CELL * (K1 * K3)
    cost = n_row + _slab.CELL_COST * cells_row + row_cost
    recs = []
    out = ["# Slab cost model: per-slab stage times on one B200 (bench workload)\n",
           "`python scripts/slab_cost_fit.py %s`; constants in use: CELL_COST = %.3f, row cost = %.0f samples\n"
           % (" ".join(map(str, Gs)), _slab.CELL_COST, row_cost),
           "| G | slab | origin rows | samples | occupied cells | axis-3 fwd+adj ms | interp fwd+adj ms | sum ms |\n|---|---|---|---|---|---|---|---|"]
    for G in Gs:
        bounds = slab_boundaries(cost, G)
        for s in range(G):
            b0, b1 = bounds[s], bounds[s + 1]
            idx = np.nonzero((rows >= b0) & (rows < b1))[0]
            k = CudaSlabKernels(Nd, Kd, (J, J, J), 1024, "single", False, (0.0, 0.0, 0.0), 1.0, None)
            nrows = b1 - b0 + J - 1
            k.make_local(om[idx], b0, nrows)
            grid = k.empty((K3, nrows, K1))
            grid.normal_()
            y = torch.randn(idx.size, dtype=torch.complex64, device="cuda")
            t_ax = timeit(lambda: (k.axis3_fwd(grid), k.axis3_adj(grid)))
            t_in = timeit(lambda: (k.interp_fwd(grid), k.interp_adj(y, grid)))
            S, C, R = float(idx.size), float(cells_row[b0:b1].sum()), float(b1 - b0)
            recs.append((S, C, R, t_ax + t_in))
            out.append("| %d | %d | %d | %d | %d | %.3f | %.3f | %.3f |" % (G, s, R, S, C, t_ax, t_in, t_ax + t_in))
            print(out[-1], flush=True)
            del k, grid, y
            torch.cuda.empty_cache()
    A = np.array([[r[0], r[1], r[2]] for r in recs])
    t = np.array([r[3] for r in recs])
    coef, *_ = np.linalg.lstsq(A, t, rcond=None)
    pred = A @ coef
    out.append("\nLeast squares over %d slabs: t = %.3e ms * samples + %.3e ms * cells + %.3e ms * rows "
               "(max relative residual %.1f %%)" % (len(recs), coef[0], coef[1], coef[2],
                                                   100 * np.max(np.abs(pred - t) / t)))
    out.append("=> CELL_COST = %.3f, row cost = %.0f samples per row of %d x %d cells (ROW_COST_PER_CELL = %.4f)"
               % (coef[1] / coef[0], coef[2] / coef[0], K1, K3, coef[2] / coef[0] / (K1 * K3)))
    for G in Gs:
        tt = [r[3] for r in recs[:G]] if G == Gs[0] else None
    txt = "\n".join(out) + "\n"
    print(txt)
    open(os.path.join(ROOT, "profiles", "r02_slab_cost_fit.md"), "w").write(txt)


if __name__ == "__main__":
    main()
