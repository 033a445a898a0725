"""Device-resident timings of BASELINE.json configs[0..4] ("also report
C1-C4 and interp-only numbers").  CUDA events, 3 warm-up + 20 timed calls each, inputs
resident in HBM.  Writes gpurun_out/configs.json and a markdown table.

    python scripts/bench_configs.py            (one B200)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mrrt.nufft_b200 import NufftBase, nufft_adj, nufft_forward  # noqa: E402
import bench  # noqa: E402


def timeit(fn, n=20, warm=3):
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


def radial2d(S, n, dtype):
    ang = np.pi * np.arange(S) / S
    r = 2 * np.pi * (np.arange(n) - n / 2) / n
    return np.stack([np.outer(np.cos(ang), r).ravel(), np.outer(np.sin(ang), r).ravel()], 1).astype(dtype)


ROWS = []


def run(name, Nd, Kd, om, J, precision, mode, coils, n=20):
    A = NufftBase(Nd=Nd, omega=om, Jd=J, Kd=Kd, precision=precision, mode=mode)
    cdt = torch.complex64 if precision == "single" else torch.complex128
    x = torch.randn((coils,) + tuple(reversed(Nd)), dtype=cdt, device="cuda")
    x = x.permute(*reversed(range(x.dim())))
    y = torch.randn((coils, A.M), dtype=cdt, device="cuda").t()
    g = torch.randn((coils, int(np.prod(Kd))), dtype=cdt, device="cuda").t()
    t_if = timeit(lambda: nufft_forward(A, g, grid_only=True), n)
    t_ia = timeit(lambda: nufft_adj(A, y, grid_only=True), n)
    t_f = timeit(lambda: A.fft(x), n)
    t_a = timeit(lambda: A.adj(y), n)
    c = 8 if precision == "single" else 16
    r = c // 2
    PN, PK, nd = int(np.prod(Nd)), int(np.prod(Kd)), len(Nd)
    # full fwd = P_N c + P_K c + 2 P_K c + [P_K c + M (ndim r + c)], adj symmetric
    pair_bytes = 2 * coils * (PN * c + 4 * PK * c + A.M * (nd * r + c))
    row = dict(config=name, M=int(A.M), coils=coils, precision=precision, mode=mode,
               interp_fwd_ms=t_if, interp_adj_ms=t_ia, full_fwd_ms=t_f, full_adj_ms=t_a,
               gpts_per_s=2 * A.M * coils / (t_f + t_a) / 1e6,
               pair_algorithmic_GB=pair_bytes / 1e9,
               hbm_GBps=pair_bytes / (t_f + t_a) / 1e6,
               fwd_kernel=A.option("last_fwd_kernel"), adj_kernel=A.option("last_adj_kernel"))
    ROWS.append(row)
    print(json.dumps(row), flush=True)
    del A, x, y, g
    torch.cuda.empty_cache()


if __name__ == "__main__":
    run("C1 2-D 256^2 radial 402x512, J=6, table", (256, 256), (512, 512),
        radial2d(402, 512, np.float32), 6, "single", "table", 1)
    rs = np.random.RandomState(1)
    om = np.clip((np.pi / 3) * rs.standard_normal((205824, 2)), -np.pi, np.pi - 1e-6)
    run("C2 2-D 256^2 random density, J=6, table", (256, 256), (512, 512), om, 6, "double", "table", 1)
    run("C2 2-D 256^2 random density, J=6, sparse", (256, 256), (512, 512), om, 6, "double", "sparse", 1)
    r2 = radial2d(201, 256, np.float64)
    kz = 2 * np.pi * (np.arange(128) - 64) / 128
    om3 = np.concatenate([np.concatenate([r2, np.full((r2.shape[0], 1), z)], 1) for z in kz], 0)
    run("C3 3-D 128^3 stack-of-stars 201x256x128, J=4", (128, 128, 128), (192, 192, 192),
        om3.astype(np.float32), 4, "single", "table", 1)
    run("C4 2-D 320^2 radial 503x640, J=6, 32 coils", (320, 320), (480, 480),
        radial2d(503, 640, np.float32), 6, "single", "table", 32)
    # 1-D (no BASELINE config; the one-thread-per-sample kernels serve it): 2^20-point grid, 8 M samples
    om1 = ((rs.rand(8 << 20, 1) * 2 - 1) * np.pi).astype(np.float32)
    run("1-D N=2^19, K=2^20, 8.4 M random samples, J=6", (1 << 19,), (1 << 20,), om1, 6, "single", "table", 1)
    om5 = bench.radial3d(bench.SPOKES, bench.NREAD)
    run("C5 3-D 256^3 radial 102944x512, J=6", bench.ND, bench.KD, om5, 6, "single", "table", 1, n=10)
    run("C5 same, double precision", bench.ND, bench.KD, om5.astype(np.float64), 6, "double", "table", 1, n=5)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(ROWS, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
    with open(os.path.join(ROOT, "gpurun_out", "configs.md"), "w") as f:
        f.write("| config | precision | M | coils | interp fwd ms | interp adj ms | fft() ms | adj() ms | "
                "G pts/s (pair) | algorithmic GB (pair) | GB/s | kernels f/a |\n|" + "---|" * 12 + "\n")
        for r in ROWS:
            f.write("| %s (%s) | %s | %d | %d | %.3f | %.3f | %.3f | %.3f | %.2f | %.3f | %.0f | %d/%d |\n" % (
                r["config"], r["mode"], r["precision"], r["M"], r["coils"], r["interp_fwd_ms"],
                r["interp_adj_ms"], r["full_fwd_ms"], r["full_adj_ms"], r["gpts_per_s"],
                r["pair_algorithmic_GB"], r["hbm_GBps"], r["fwd_kernel"], r["adj_kernel"]))
