"""Toeplitz normal operator (ToeplitzNorm) against NufftBase.norm = adj(fft(x)) on the bench
workload (BASELINE configs[4]) and on configs[3] (2-D 320^2, 32 coils).  One JSON line each."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mrrt.nufft_b200 import NufftBase, ToeplitzNorm  # noqa: E402
from sense_bench import radial2d, timeit  # noqa: E402


def case(name, Nd, Kd, om, reps, n=10):
    dev = torch.device("cuda", 0)
    A = NufftBase(Nd=Nd, omega=om, Jd=6, Kd=Kd, precision="single")
    t0 = time.perf_counter()
    T = ToeplitzNorm(A)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    shape = tuple(reversed(Nd)) if reps == 1 else (reps,) + tuple(reversed(Nd))
    x = torch.randn(shape, dtype=torch.complex64, device=dev)
    x = x.permute(*reversed(range(x.dim())))
    yt, ya = T.norm(x), A.norm(x)
    r = {"case": name, "M": int(A.M), "reps": reps, "toeplitz_setup_s": t_setup,
         "toeplitz_norm_ms": timeit(lambda: T.norm(x), n), "nufft_norm_ms": timeit(lambda: A.norm(x), n),
         "rel_l2_toeplitz_vs_nufft_norm": float((yt - ya).norm() / ya.norm())}
    print(json.dumps(r), flush=True)


if __name__ == "__main__":
    case("C4 2-D 320^2 32 coils", (320, 320), (480, 480), radial2d(503, 640), 32, n=30)
    case("C5 3-D 256^3 (bench workload)", bench.ND, bench.KD, bench.radial3d(bench.SPOKES, bench.NREAD), 1)
