"""Run-to-run spread of the one-RED-per-tap fallback adjoint (force_generic) in float32 on the
mid-size 3-D radial case of tests/test_gpu_parity.py: distance to the reference's float32 result
and to the float64 twin, ten launches (the order of the float atomics differs between launches)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from golden_util import rel_l2  # noqa: E402
from test_gpu_parity import _radial3d  # noqa: E402
from oracle import nufft_oracle as orc  # noqa: E402
from mrrt.nufft_b200 import NufftBase  # noqa: E402

Nd, Kd = (32, 32, 32), (48, 48, 48)
om = _radial3d(700, 64).astype(np.float32)
O = orc.OracleNufft(Nd=Nd, omega=om, Jd=6, Kd=Kd, precision="single",
                    engine="reference" if orc.have_reference_engine() else "port")
T = orc.float64_twin(O)
rs = np.random.RandomState(0)
x = (rs.standard_normal(Nd) + 1j * rs.standard_normal(Nd)).astype(np.complex64)
yo = O.fft(x)
xo, xt = O.adj(yo), T.adj(yo.astype(np.complex128))
print("reference vs twin %.3g" % rel_l2(xo, xt))
for name, opts in (("generic", {"force_generic": 1}), ("auto", {})):
    A = NufftBase(Nd=Nd, omega=om, Jd=6, Kd=Kd, precision="single", options=opts)
    for k in range(10):
        xa = A.adj(yo)
        print("%s run %d: vs reference %.3g, vs twin %.3g" % (name, k, rel_l2(xa, xo), rel_l2(xa, xt)))
