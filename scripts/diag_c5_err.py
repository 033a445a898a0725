"""Diagnostic: margins of the float32 parity checks on the bench-workload subsample."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
from oracle import nufft_oracle as orc
from golden_util import rel_l2
from mrrt.nufft_b200 import NufftBase

idx = np.arange(0, bench.SPOKES, int(sys.argv[1]) if len(sys.argv) > 1 else 128)
om = np.concatenate([bench.radial3d(bench.SPOKES, bench.NREAD, int(s), int(s) + 1) for s in idx], 0)
A = NufftBase(Nd=bench.ND, omega=om, Jd=bench.JD, Kd=bench.KD, precision="single")
O = orc.OracleNufft(Nd=bench.ND, omega=om, Jd=bench.JD, Kd=bench.KD, precision="single", engine="reference")
x = bench.image()
yo = O.fft(x)
xo = O.adj(yo)
for rep in range(3):
    print("run %d: fft rel-L2 %.3g   adj rel-L2 %.3g" % (rep, rel_l2(A.fft(x), yo), rel_l2(A.adj(yo), xo)))
B = NufftBase(Nd=bench.ND, omega=om, Jd=bench.JD, Kd=bench.KD, precision="single", options={"pruned_fft": 0})
print("unpruned FFT: fft %.3g adj %.3g" % (rel_l2(B.fft(x), yo), rel_l2(B.adj(yo), xo)))
# where does the adjoint difference come from?  float64 oracle on the SAME float32 tm
O64 = orc.OracleNufft(Nd=bench.ND, omega=om.astype(np.float64), Jd=bench.JD, Kd=bench.KD,
                      precision="double", engine="reference")
O64.tm = O.tm.astype(np.float64)
O64.phase_before = O.phase_before.astype(np.complex128)   # same (float32-rounded) phases
O64.phase_after = O.phase_after.astype(np.complex128)
x64 = O64.adj(yo.astype(np.complex128))
print("adjoint vs float64 evaluation of the same operator: reference float32 path %.3g, CUDA path %.3g"
      % (rel_l2(xo, x64), rel_l2(A.adj(yo), x64)))
