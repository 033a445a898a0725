"""Device timing of the 2-D configurations (BASELINE configs[0], [1], [3])."""
import sys, time
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, ".."); sys.path.insert(0, "scripts")
from mrrt.nufft_b200 import NufftBase, nufft_adj, nufft_forward
from quick_bench import timeit


def radial2d(S, n, dtype):
    ang = np.pi * np.arange(S) / S
    r = 2 * np.pi * (np.arange(n) - n / 2) / n
    return np.stack([np.outer(np.cos(ang), r).ravel(), np.outer(np.sin(ang), r).ravel()], 1).astype(dtype)


def run(name, Nd, Kd, om, J, precision, mode, coils, opts={}):
    A = NufftBase(Nd=Nd, omega=om, Jd=J, Kd=Kd, precision=precision, mode=mode, options=opts)
    cdt = torch.complex64 if precision == "single" else torch.complex128
    x = torch.randn((coils,) + tuple(reversed(Nd)), dtype=cdt, device="cuda")
    x = x.permute(*reversed(range(x.dim())))
    y = torch.randn((coils, A.M), dtype=cdt, device="cuda").t()
    g = torch.randn((coils, int(np.prod(Kd))), dtype=cdt, device="cuda").t()
    t_if = timeit(lambda: nufft_forward(A, g, grid_only=True), n=20)
    t_ia = timeit(lambda: nufft_adj(A, y, grid_only=True), n=20)
    t_f = timeit(lambda: A.fft(x), n=20)
    t_a = timeit(lambda: A.adj(y), n=20)
    print("%-44s M=%7d coils=%2d | interp fwd %7.3f adj %7.3f ms | full fwd %7.3f adj %7.3f ms | %.2f Gpt/s k=%d/%d" % (
        name, A.M, coils, t_if, t_ia, t_f, t_a, 2 * A.M * coils / (t_f + t_a) / 1e6,
        A.option("last_fwd_kernel"), A.option("last_adj_kernel")), flush=True)


if __name__ == "__main__":
    sys.path.insert(0, "scripts")
    run("C1 2D 256^2 radial 402x512 c64 table", (256, 256), (512, 512), radial2d(402, 512, np.float32), 6, "single", "table", 1)
    rs = np.random.RandomState(1)
    om = np.clip((np.pi / 3) * rs.standard_normal((205824, 2)), -np.pi, np.pi - 1e-6)
    run("C2 2D 256^2 random-density c128 table", (256, 256), (512, 512), om, 6, "double", "table", 1)
    run("C2 2D 256^2 random-density c128 sparse", (256, 256), (512, 512), om, 6, "double", "sparse", 1)
    run("C4 2D 320^2 radial 503x640 c64 32 coils", (320, 320), (480, 480), radial2d(503, 640, np.float32), 6, "single", "table", 32)
    run("C4 same, generic", (320, 320), (480, 480), radial2d(503, 640, np.float32), 6, "single", "table", 32, {"force_generic": 1})
    # C3: 3-D 128^3 stack-of-stars J=4 Kd=192^3
    r2 = radial2d(201, 256, np.float64)
    kz = 2 * np.pi * (np.arange(128) - 64) / 128
    om3 = np.concatenate([np.concatenate([r2, np.full((r2.shape[0], 1), z)], 1) for z in kz], 0).astype(np.float32)
    run("C3 3D 128^3 stack-of-stars J=4 c64", (128, 128, 128), (192, 192, 192), om3, 4, "single", "table", 1)
