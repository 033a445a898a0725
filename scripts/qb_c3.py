"""C3 (3-D 128^3 stack-of-stars, J=4) forward interpolation with pairing off / auto / on."""
import sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
from mrrt.nufft_b200 import NufftBase, nufft_forward
from bench_configs import radial2d, timeit
r2 = radial2d(201, 256, np.float64)
kz = 2 * np.pi * (np.arange(128) - 64) / 128
om3 = np.concatenate([np.concatenate([r2, np.full((r2.shape[0], 1), z)], 1) for z in kz], 0).astype(np.float32)
for J in (4, 6):
    for fp in (0, 1, 2):
        A = NufftBase(Nd=(128,) * 3, omega=om3, Jd=J, Kd=(192,) * 3, precision="single", options={"fwd_pair": fp})
        g = torch.randn((1, 192 ** 3), dtype=torch.complex64, device="cuda").t()
        t = timeit(lambda: nufft_forward(A, g, grid_only=True), 20)
        print("C3 J=%d fwd_pair=%d: interp fwd %.3f ms, slots/M %.3f" % (J, fp, t, A.option("n_slots") / A.M), flush=True)
        del A
