"""A/B timing of adjoint-kernel plan options on the bench workload (interpolation stage only).
usage: python scripts/ab_adj.py '{"win_facew": 1}' '{"slide_pts": 512}' ...   (baseline {} always runs first)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mrrt.nufft_b200 import NufftBase, nufft_adj, nufft_forward  # noqa: E402


def timeit(f, n=10, warm=3):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        f()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


if __name__ == "__main__":
    small = os.environ.get("AB_SMALL") == "1"
    JJ = int(os.environ.get("AB_J", "6"))
    if os.environ.get("AB_CASE") == "c3":
        # BASELINE configs[2]: 3-D 128^3 stack-of-stars 201 x 256 x 128, Kd = 192^3
        from sense_bench import stack_of_stars
        om, Nd, Kd = stack_of_stars(201, 256, 128), (128,) * 3, (192,) * 3
    else:
        om = bench.radial3d(bench.SPOKES // (8 if small else 1), bench.NREAD // (2 if small else 1))
        Nd, Kd = ((128,) * 3, (192,) * 3) if small else (bench.ND, bench.KD)
    variants = [{}] + [json.loads(a) for a in sys.argv[1:]]
    y = g = ref_a = ref_f = None
    prec = os.environ.get("AB_PRECISION", "single")
    cdt = torch.complex64 if prec == "single" else torch.complex128
    for opts in variants:
        A = NufftBase(Nd=Nd, omega=om, Jd=JJ, Kd=Kd, precision=prec, options=opts)
        if y is None:
            y = torch.randn(A.M, dtype=cdt, device="cuda")
            g = torch.randn(int(np.prod(Kd)), dtype=cdt, device="cuda")
        ga = nufft_adj(A, y, grid_only=True)
        gf = nufft_forward(A, g, grid_only=True)
        if ref_a is None:
            ref_a, ref_f = ga.clone(), gf.clone()
        xim = torch.randn(tuple(reversed(Nd)), dtype=cdt, device="cuda").permute(2, 1, 0)
        r = {"opts": opts, "fft_full_ms": timeit(lambda: A.fft(xim)), "adj_full_ms": timeit(lambda: A.adj(y)),
             "adj_ms": timeit(lambda: nufft_adj(A, y, grid_only=True)),
             "fwd_ms": timeit(lambda: nufft_forward(A, g, grid_only=True)),
             "adj_rel_vs_base": float((ga - ref_a).norm() / ref_a.norm()),
             "fwd_rel_vs_base": float((gf - ref_f).norm() / ref_f.norm()),
             "kernels": [A.option("last_fwd_kernel"), A.option("last_adj_kernel")]}
        print(json.dumps(r), flush=True)
        del A
        torch.cuda.empty_cache()
