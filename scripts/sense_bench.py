"""Coil-sensitivity fusion (SenseNufft) against the unfused caller sequence, BASELINE
configs[3] geometry (2-D 320^2, 503x640 radial, 32 coils, Kd=480^2, J=6, complex64) and a
3-D 128^3 8-coil case.  Prints one JSON line per case."""
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrrt.nufft_b200 import NufftBase, SenseNufft  # noqa: E402


def radial2d(S, n):
    ang = np.pi * np.arange(S) / S
    r = 2 * np.pi * (np.arange(n) - n / 2) / n
    return np.stack([np.outer(np.cos(ang), r).ravel(), np.outer(np.sin(ang), r).ravel()], 1).astype(np.float32)


def stack_of_stars(S, n, P):
    om2 = radial2d(S, n)
    kz = 2 * np.pi * (np.arange(P) - P // 2) / P
    return np.concatenate([np.concatenate([om2, np.full((om2.shape[0], 1), z, np.float32)], 1)
                           for z in kz], 0).astype(np.float32)


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def case(name, Nd, Kd, om, nc, J=6):
    rs = np.random.RandomState(0)
    dev = torch.device("cuda", 0)
    smaps = torch.from_numpy((rs.standard_normal(Nd + (nc,)) + 1j * rs.standard_normal(Nd + (nc,)))
                             .astype(np.complex64)).to(dev)
    smaps = smaps.permute(*reversed(range(smaps.dim()))).contiguous().permute(*reversed(range(smaps.dim())))
    S = SenseNufft(Nd=Nd, omega=om, smaps=smaps, Jd=J, Kd=Kd, precision="single")
    A = S.op
    x = torch.from_numpy((rs.standard_normal(Nd) + 1j * rs.standard_normal(Nd)).astype(np.complex64)).to(dev)
    x = x.permute(*reversed(range(x.dim()))).contiguous().permute(*reversed(range(x.dim())))
    k = S.fft(x)
    csm = smaps.conj()

    def unf_fwd():
        return A.fft(x[..., None] * smaps)

    def unf_adj():
        return (csm * A.adj(k)).sum(-1)

    r = {"case": name, "M": int(A.M), "ncoil": nc,
         "fused_fft_ms": timeit(lambda: S.fft(x)), "unfused_fft_ms": timeit(unf_fwd),
         "fused_adj_ms": timeit(lambda: S.adj(k)), "unfused_adj_ms": timeit(unf_adj),
         "fused_norm_ms": timeit(lambda: S.norm(x)),
         "unfused_norm_ms": timeit(lambda: (csm * A.adj(A.fft(x[..., None] * smaps))).sum(-1))}
    ku = unf_fwd()
    r["fft_rel_l2_fused_vs_unfused"] = float((k - ku).norm() / ku.norm())
    xa, xu = S.adj(k), unf_adj()
    r["adj_rel_l2_fused_vs_unfused"] = float((xa - xu).norm() / xu.norm())
    print(json.dumps(r), flush=True)


if __name__ == "__main__":
    case("C4 2-D 320^2 32 coils", (320, 320), (480, 480), radial2d(503, 640), 32)
    case("3-D 128^3 stack-of-stars 8 coils J=4", (128, 128, 128), (192, 192, 192),
         stack_of_stars(201, 256, 128), 8, J=4)
