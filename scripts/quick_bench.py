"""Ad-hoc device timing of the interpolation stages (not the contract bench)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from mrrt.nufft_b200 import NufftBase, nufft_adj, nufft_forward  # noqa: E402


def radial3d(S, n):
    s = np.arange(S, dtype=np.float64)
    z = 1 - (2 * s + 1) / S
    phi = s * np.pi * (3 - np.sqrt(5))
    rxy = np.sqrt(1 - z * z)
    d = np.stack([rxy * np.cos(phi), rxy * np.sin(phi), z], 1).astype(np.float32)
    r = (2 * np.pi * (np.arange(n) - n // 2) / n).astype(np.float32)
    return (d[:, None, :] * r[None, :, None]).reshape(-1, 3)


def timeit(f, n=5, warm=2):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        f()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def run(name, Nd, Kd, om, J, opts, precision="single"):
    t0 = time.time()
    A = NufftBase(Nd=Nd, omega=om, Jd=J, Kd=Kd, precision=precision, options=opts)
    torch.cuda.synchronize()
    tplan = time.time() - t0
    PK = int(np.prod(Kd))
    cdt = torch.complex64 if precision == "single" else torch.complex128
    g = torch.randn(PK, dtype=cdt, device="cuda")
    y = torch.randn(A.M, dtype=cdt, device="cuda")
    x = torch.randn(tuple(reversed(Nd)), dtype=cdt, device="cuda").permute(2, 1, 0)
    t_if = timeit(lambda: nufft_forward(A, g, grid_only=True))
    t_ia = timeit(lambda: nufft_adj(A, y, grid_only=True))
    t_f = timeit(lambda: A.fft(x))
    t_a = timeit(lambda: A.adj(y))
    print("%-28s M=%9d plan %.1fs | interp fwd %8.3f ms adj %8.3f ms | full fwd %8.3f adj %8.3f ms | %.2f Gpt/s | items %d k=%d/%d" % (
        name, A.M, tplan, t_if, t_ia, t_f, t_a, 2 * A.M / (t_f + t_a) / 1e6,
        A.option("n_items"), A.option("last_fwd_kernel"), A.option("last_adj_kernel")), flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "small":
        om = radial3d(12868, 256)
        for opts in ({}, {"force_generic": 1}, {"use_tma": 0}):
            run("128^3 K192 J6 " + str(opts), (128,) * 3, (192,) * 3, om, 6, opts)
    else:
        om = radial3d(102944, 512)
        for opts in ({}, {"use_tma": 0}, {"chunk": 4096}, {"slide_pts": 1024}, {"tile1": 32, "tile2": 8, "tile3": 8}):
            run("C5 256^3 K384 J6 " + str(opts), (256,) * 3, (384,) * 3, om, 6, opts)
        run("C5 generic", (256,) * 3, (384,) * 3, om, 6, {"force_generic": 1})
