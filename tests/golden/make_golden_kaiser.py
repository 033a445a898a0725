"""Golden values of the Kaiser-Bessel functions and kernel objects, produced by the REAL
reference package in this container (same recipe as make_golden.py):

    oracle/build_reference_scratch.sh /tmp/refbuild
    PYTHONPATH=/tmp/refbuild python tests/golden/make_golden_kaiser.py

Parameters are the reference's own test parameters (tests/test_kaiser.py:18-88,
tests/test_kernels.py:17-58) plus the BeattyKernel shapes of the BASELINE configs."""
import os
import warnings

import numpy as np

warnings.simplefilter("ignore")

from mrrt.nufft._kaiser_bessel import kaiser_bessel, kaiser_bessel_ft  # noqa: E402
from mrrt.nufft._kernels import BeattyKernel  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
out = {}
# tests/test_kaiser.py:18-31
J, alpha = 8, 2.34 * 8
x = np.linspace(-(J + 1) / 2.0, (J + 1) / 2.0, 1001)
out["kb_x"] = x
for m in (-4, 0, 2, 7):
    out["kb_J8_m%d" % m] = kaiser_bessel(x, J, alpha, m)
# tests/test_kaiser.py:61-88
J, alpha, N = 5, 6.8, 2 ** 10
xx = np.arange(-N / 2.0, N / 2.0) / float(N) * (J + 3) / 2.0
uu = 1.5 * np.linspace(-1, 1, 201)
out["ft_u"] = uu
for m in (-2, 0, 2, 7):
    out["ft_J5_m%d" % m] = kaiser_bessel_ft(uu, J, alpha, m, 1)
    out["kbx_J5_m%d" % m] = kaiser_bessel(xx, J, alpha, m)
out["kbx_x"] = xx
# BeattyKernel parameters and values (tests/test_kernels.py + BASELINE shapes)
for tag, (shape, grid, os_grid) in {
    "c1": ((6, 6), (256, 256), (512, 512)),
    "c3": ((4, 4, 4), (128, 128, 128), (192, 192, 192)),
    "c5": ((6, 6, 6), (256, 256, 256), (384, 384, 384)),
    "odd": ((3, 4), (64, 64), (128, 128)),
    "t": ((4, 4), (24, 16), (32, 32)),
}.items():
    k = BeattyKernel(shape=shape, grid_shape=grid, os_grid_shape=os_grid)
    out["beatty_%s_alpha" % tag] = np.asarray(k.alpha, dtype=np.float64)
    out["beatty_%s_m" % tag] = np.asarray(k.m, dtype=np.float64)
    for d in range(len(shape)):
        xs = np.linspace(-shape[d] / 2 - 0.5, shape[d] / 2 + 0.5, 301)
        out["beatty_%s_k%d" % (tag, d)] = k.kernels[d](xs)
np.savez_compressed(os.path.join(HERE, "extra", "kaiser.npz"), **out)
print("wrote", len(out), "arrays")
