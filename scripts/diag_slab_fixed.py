"""Diagnostic: staged slab transforms vs the single-GPU operator on a fixed-schedule geometry."""
import os, sys
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from golden_util import rel_l2
from test_slab_gpu import _radial3d
from oracle import nufft_oracle as orc
from mrrt.nufft_b200 import NufftBase
from mrrt.nufft_b200._slab import CudaSlabKernels, _pieces, row_statistics, slab_boundaries

precision = "single"; G = 3
Nd, Kd, J = (80, 100, 70), (128, 192, 128), 6
n_shift = (3.0, 0.0, 1.5)
rdt = np.dtype(np.float32)
om = _radial3d(600, 64).astype(rdt)
rs = np.random.RandomState(2)
om[:200] = ((rs.rand(200, 3) * 2 - 1) * np.pi).astype(rdt)
om[200:230, 1] = np.pi - 1e-3
M = om.shape[0]
O = orc.OracleNufft(Nd=Nd, omega=om, Jd=J, Kd=Kd, precision=precision, n_shift=n_shift)
x = (rs.standard_normal(Nd) + 1j * rs.standard_normal(Nd)).astype(np.complex64)
yo = O.fft(x)
for opts in ({}, {"own_fft12": 0}, {"own_fft3": 2}, {"own_fft3": 0}):
    A = NufftBase(Nd=Nd, omega=om, Jd=J, Kd=Kd, precision=precision, n_shift=n_shift, options=opts)
    print(opts, "A vs oracle %.3g" % rel_l2(A.fft(x), yo), "inplane", A.option("inplane_own"))
A = NufftBase(Nd=Nd, omega=om, Jd=J, Kd=Kd, precision=precision, n_shift=n_shift)
K1, K2, K3 = Kd; N3 = Nd[2]
rows, n_row, cells_row = row_statistics(om, (J, J, J), Kd, rdt, torch.device("cuda"))
bounds = slab_boundaries(n_row + 1.38 * cells_row + 20.0, G)
for kopts in ({}, {'own_fft12': 0}, {'own_fft3': 2}, {'own_fft3': 0}, {'fwd_pair': 0}, {'force_generic': 1}):
    print('--- slab kernels options', kopts)
    ranks = []
    for s in range(G):
        k = CudaSlabKernels(Nd, Kd, (J, J, J), 1024, precision, False, n_shift, 1.0, None, kopts)
        idx = np.nonzero((rows >= bounds[s]) & (rows < bounds[s + 1]))[0]
        k.make_local(om[idx], bounds[s], bounds[s + 1] - bounds[s] + J - 1)
        ranks.append((k, idx, bounds[s], bounds[s + 1] - bounds[s] + J - 1))
    k0 = ranks[0][0]
    xp = k0.to_device(np.ascontiguousarray(x.transpose(2, 1, 0)))
    Apl = torch.cat([k0.planes_fwd(xp[:10].contiguous(), 0), k0.planes_fwd(xp[10:].contiguous(), 10)], 0)
    Apl1 = k0.planes_fwd(xp.contiguous(), 0)
    print("planes split vs whole", float((Apl - Apl1).norm() / Apl1.norm()))
    yy = np.zeros(M, dtype=np.complex64)
    for k, idx, row0, nrows in ranks:
        grid = k.empty((K3, nrows, K1)); grid.zero_()
        for glo, llo, n in _pieces(row0, nrows, K2):
            grid[:N3, llo:llo + n] = Apl[:, glo:glo + n]
        k.axis3_fwd(grid)
        yy[idx] = k.interp_fwd(grid).cpu().numpy()
        print("rank rows", row0, nrows, "axis3_fused", k.axis3_fused, "err vs A on its samples %.3g" % rel_l2(yy[idx], A.fft(x)[idx]))
    print("staged vs A %.3g, staged vs oracle %.3g" % (rel_l2(yy, A.fft(x)), rel_l2(yy, yo)))
    
