"""Diagnostic: float32 adjoint error budget of the CUDA kernels vs the CPU oracle."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import nufft_oracle as orc
from test_gpu_parity import _radial3d
from golden_util import rel_l2
from mrrt.nufft_b200 import NufftBase, nufft_adj, nufft_forward

Nd, Kd = (32, 32, 32), (48, 48, 48)
om = _radial3d(700, 64).astype(np.float32)
O32 = orc.OracleNufft(Nd=Nd, omega=om, Jd=6, Kd=Kd, precision="single", engine="reference")
O64 = orc.OracleNufft(Nd=Nd, omega=om.astype(np.float64), Jd=6, Kd=Kd, precision="double", engine="reference")
O64.tm = O32.tm.astype(np.float64)
rs = np.random.RandomState(0)
x = (rs.standard_normal(Nd) + 1j * rs.standard_normal(Nd)).astype(np.complex64)
yo = O32.fft(x)
g32 = O32.adj(yo, grid_only=True); g64 = O64.adj(yo.astype(np.complex128), grid_only=True)
a32 = O32.adj(yo); a64 = O64.adj(yo.astype(np.complex128))
print("oracle32 vs oracle64: grid %.3g full %.3g" % (rel_l2(g32, g64), rel_l2(a32, a64)))
for name, opts in (("auto", {}), ("generic", {"force_generic": 1})):
    A = NufftBase(Nd=Nd, omega=om, Jd=6, Kd=Kd, precision="single", options=opts)
    for rep in range(3):
        g = nufft_adj(A, yo, grid_only=True).cpu().numpy()
        a = A.adj(yo)
        print("%-8s run %d: grid vs o32 %.3g vs o64 %.3g | full vs o32 %.3g vs o64 %.3g" % (
            name, rep, rel_l2(g, g32), rel_l2(g, g64), rel_l2(a, a32), rel_l2(a, a64)))
    f = A.fft(x)
    print("%-8s fwd vs o32 %.3g vs o64 %.3g (o32 vs o64 %.3g)" % (name, rel_l2(f, yo), rel_l2(f, O64.fft(x.astype(np.complex128))), rel_l2(yo, O64.fft(x.astype(np.complex128)))))
