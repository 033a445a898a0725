"""The slab-distributed operator on real GPUs.

* ``test_slab_plans_one_gpu``: the C ABI's slab plans and staged transforms (b2n_planes_fwd/adj,
  b2n_axis3_fwd/adj, options slab_kglobal2 / slab_origin2) with G simulated ranks on ONE GPU --
  the all-to-all is replaced by local copies -- against the single-plan operator and the oracle.
* ``test_slab_sharded_nccl``: ``SlabShardedNufft`` itself, 2 ranks over NCCL, against the
  single-GPU result (needs 2 GPUs; skipped otherwise)."""
import os
import socket

import numpy as np
import pytest

from golden_util import TOL, rel_l2

pytestmark = pytest.mark.gpu


def _radial3d(S, n):
    s = np.arange(S)
    z = 1 - (2 * s + 1) / S
    phi = s * np.pi * (3 - np.sqrt(5))
    rxy = np.sqrt(1 - z * z)
    d = np.stack([rxy * np.cos(phi), rxy * np.sin(phi), z], 1)
    r = 2 * np.pi * (np.arange(n) - n // 2) / n
    return (d[:, None, :] * r[None, :, None]).reshape(-1, 3)


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("G,geom", [(1, "small"), (2, "small"), (5, "small"), (3, "fixed")])
def test_slab_plans_one_gpu(G, geom, precision):
    import torch
    from oracle import nufft_oracle as orc
    from mrrt.nufft_b200 import NufftBase
    from mrrt.nufft_b200._slab import CudaSlabKernels, _pieces, row_statistics, slab_boundaries

    # "fixed": every Kd has a compile-time FFT schedule, so the plane stage and the axis-3 stage run
    # the own line-FFT passes (scale/pad and crop fused) instead of cuFFT
    Nd, Kd, J = ((32, 28, 24), (48, 44, 36), 6) if geom == "small" else ((80, 100, 70), (128, 192, 128), 6)
    # (fixed: n_shift = Nd / 2 makes phase_after exactly 1.  With |angle| ~ 400 the float32 dot
    # product behind phase_after differs by an ulp (3e-5 rad) for some rows between a BLAS call
    # on all samples and one on a rank's subset -- host BLAS blocking, 3e-6 in the result, still
    # inside the 1e-5 parity with the oracle but not inside the tol / 4 self-consistency below)
    n_shift = (3.0, 0.0, 1.5) if geom == "small" else tuple(n / 2.0 for n in Nd)
    rdt = np.dtype(np.float32 if precision == "single" else np.float64)
    om = _radial3d(600, 64).astype(rdt)
    rs = np.random.RandomState(2)
    om[:200] = ((rs.rand(200, 3) * 2 - 1) * np.pi).astype(rdt)      # fill the corners too
    om[200:230, 1] = np.pi - 1e-3                                    # windows across the seam
    M = om.shape[0]
    A = NufftBase(Nd=Nd, omega=om, Jd=J, Kd=Kd, precision=precision, n_shift=n_shift)
    O = orc.OracleNufft(Nd=Nd, omega=om, Jd=J, Kd=Kd, precision=precision, n_shift=n_shift)
    cdt = A._cplx_dtype
    x = (rs.standard_normal(Nd) + 1j * rs.standard_normal(Nd)).astype(cdt)
    y = (rs.standard_normal(M) + 1j * rs.standard_normal(M)).astype(cdt)
    K1, K2, K3 = Kd
    N3 = Nd[2]
    rows, n_row, cells_row = row_statistics(om, (J, J, J), Kd, rdt, torch.device("cuda"))
    bounds = slab_boundaries(n_row + 1.38 * cells_row + 20.0, G) if G > 1 else [0, K2]
    halo = J - 1 if G > 1 else 0
    ranks = []
    for s in range(G):
        k = CudaSlabKernels(Nd, Kd, (J, J, J), 1024, precision, False, n_shift, 1.0, None)
        idx = np.nonzero((rows >= bounds[s]) & (rows < bounds[s + 1]))[0]
        k.make_local(om[idx], bounds[s], bounds[s + 1] - bounds[s] + halo)
        ranks.append((k, idx, bounds[s], bounds[s + 1] - bounds[s] + halo))
    k0 = ranks[0][0]
    # ---- forward: plane stage (all planes at once), rows to every slab, axis 3, interpolation
    xp = k0.to_device(np.ascontiguousarray(x.transpose(2, 1, 0)))
    Apl = torch.cat([k0.planes_fwd(xp[:10].contiguous(), 0), k0.planes_fwd(xp[10:].contiguous(), 10)], 0)
    yy = np.zeros(M, dtype=cdt)
    for k, idx, row0, nrows in ranks:
        grid = k.empty((K3, nrows, K1))
        grid.zero_()
        for glo, llo, n in _pieces(row0, nrows, K2):
            grid[:N3, llo:llo + n] = Apl[:, glo:glo + n]
        k.axis3_fwd(grid)
        yy[idx] = k.interp_fwd(grid).cpu().numpy()
        if G > 1 and idx.size:
            assert k.option("last_fwd_kernel") == 1
    tol = TOL[precision]
    assert rel_l2(yy, A.fft(x)) <= tol / 4
    assert rel_l2(yy, O.fft(x)) <= tol
    # ---- adjoint: gridding per slab, axis 3, halo rows summed, plane stage
    B = k0.empty((N3, K2, K1))
    B.zero_()
    for k, idx, row0, nrows in ranks:
        grid = k.empty((K3, nrows, K1))
        k.interp_adj(k.to_device(y[idx]), grid)
        if G > 1 and idx.size:
            assert k.option("last_adj_kernel") in (3, 5)
        k.axis3_adj(grid)
        for glo, llo, n in _pieces(row0, nrows, K2):
            B[:, glo:glo + n] += grid[:N3, llo:llo + n]
    xa = torch.cat([k0.planes_adj(B[:7].contiguous(), 0), k0.planes_adj(B[7:].contiguous(), 7)], 0)
    xa = xa.permute(2, 1, 0).cpu().numpy()
    assert rel_l2(xa, A.adj(y)) <= tol / 2
    assert rel_l2(xa, O.adj(y)) <= tol
    # a sample outside the slab's rows is rejected loudly
    if G > 1:
        k, idx, row0, nrows = ranks[0]
        other = ranks[1][1]
        bad = CudaSlabKernels(Nd, Kd, (J, J, J), 1024, precision, False, n_shift, 1.0, None)
        far = other[np.argmax((rows[other] - row0) % K2)]
        if (rows[far] - row0) % K2 + J > nrows:
            with pytest.raises(ValueError):
                bad.make_local(om[[far]], row0, nrows)


def test_row_statistics_on_device_match_host():
    """The device-side row computation (torch) is bit-identical to the host formula and to the
    plan's own window origins, also for coordinates within an ulp of a cell boundary."""
    import torch
    from mrrt.nufft_b200._slab import row_statistics, window_rows

    rs = np.random.RandomState(5)
    Kd, Jd = (384, 384, 384), (6, 6, 6)
    om = ((rs.rand(2000000, 3) * 2 - 1) * np.pi).astype(np.float32)
    k = rs.randint(-190, 190, 200000)                      # on cell boundaries, +- one ulp
    edge = (k * np.float32(2 * np.pi / 384)).astype(np.float32)
    om[:200000, 1] = np.nextafter(edge, np.float32(np.where(rs.rand(200000) < 0.5, -10, 10)))
    om[200000:400000, 1] = edge
    rows, n, cells = row_statistics(om, Jd, Kd, np.dtype(np.float32), torch.device("cuda"))
    assert np.array_equal(rows, window_rows(om[:, 1], 6, 384, np.dtype(np.float32)))
    assert n.sum() == om.shape[0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _nccl_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        _nccl_body(rank, q)
    except Exception:                       # report instead of leaving the parent waiting
        import traceback

        q.put({"rank": rank, "error": traceback.format_exc()})
    finally:
        dist.destroy_process_group()


def _nccl_body(rank, q):
    if True:
        from mrrt.nufft_b200 import NufftBase, SampleShardedNufft, SlabShardedNufft

        Nd, Kd, J = (48, 40, 32), (72, 60, 48), 6
        om = _radial3d(3000, 96).astype(np.float32)
        rs = np.random.RandomState(0)
        x = (rs.standard_normal(Nd) + 1j * rs.standard_normal(Nd)).astype(np.complex64)
        y = (rs.standard_normal(om.shape[0]) + 1j * rs.standard_normal(om.shape[0])).astype(np.complex64)
        A = NufftBase(Nd=Nd, omega=om, Jd=J, Kd=Kd, precision="single")
        S = SlabShardedNufft(Nd, om, Jd=J, Kd=Kd, precision="single")
        res = {"rank": rank, "M": S.M}
        y1 = A.fft(x)
        res["fwd"] = rel_l2(S.fft(x), y1[S.index])
        res["gather"] = rel_l2(S.gather_samples(S.fft(x)), y1)
        x1 = A.adj(y)
        res["adj"] = rel_l2(S.adj(y[S.index]), x1)
        res["adj_planes"] = rel_l2(S.adj(y[S.index], planes=True), x1[:, :, S.z0:S.z1])
        T = SampleShardedNufft(Nd, om, Jd=J, Kd=Kd, precision="single")
        res["sample_fwd"] = rel_l2(T.fft(x), y1[T.lo:T.hi])
        res["sample_adj"] = rel_l2(T.adj(y[T.lo:T.hi]), x1)
        q.put(res)


def test_slab_sharded_nccl():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    for r in out:
        assert "error" not in r, r["error"]
    assert all(p.exitcode == 0 for p in procs)
    assert sum(r["M"] for r in out) == 3000 * 96
    for r in out:
        for key in ("fwd", "gather", "adj", "adj_planes", "sample_fwd", "sample_adj"):
            assert r[key] <= 2.5e-6, (key, r)


def _p2p_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from mrrt.nufft_b200 import NufftBase, SlabShardedNufft

        Nd, Kd, J = (48, 40, 32), (72, 60, 48), 6
        om = _radial3d(3000, 96).astype(np.float32)
        rs = np.random.RandomState(0)
        x = (rs.standard_normal(Nd) + 1j * rs.standard_normal(Nd)).astype(np.complex64)
        y = (rs.standard_normal(om.shape[0]) + 1j * rs.standard_normal(om.shape[0])).astype(np.complex64)
        A = NufftBase(Nd=Nd, omega=om, Jd=J, Kd=Kd, precision="single")
        S = SlabShardedNufft(Nd, om, Jd=J, Kd=Kd, precision="single", exchange="auto")
        res = {"rank": rank, "exchange": S.exchange}
        y1, x1 = A.fft(x), A.adj(y)
        for rep in range(3):                      # repeated calls: the barriers protect grid reuse
            res["fwd%d" % rep] = rel_l2(S.fft(x), y1[S.index])
            res["adj%d" % rep] = rel_l2(S.adj(y[S.index]), x1)
        res["fwd_twice"] = rel_l2(S.fft(x), y1[S.index]) + rel_l2(S.fft(x), y1[S.index])
        q.put(res)
    except Exception:
        import traceback

        q.put({"rank": rank, "error": traceback.format_exc()})
    finally:
        dist.destroy_process_group()


def test_slab_sharded_peer_memory_exchange():
    """exchange="auto": rows travel by direct stores / loads on symmetric (peer-mapped) memory
    when torch's symmetric memory is available on the box, else over NCCL -- same results."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_p2p_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    for r in out:
        assert "error" not in r, r["error"]
    print("exchange used:", [r["exchange"] for r in out])
    assert out[0]["exchange"] == out[1]["exchange"]
    for r in out:
        for k, v in r.items():
            if k.startswith(("fwd", "adj")):
                assert v <= 5e-6, (k, r)


@pytest.mark.parametrize("precision", ["single", "double"])
def test_slab_exchange_kernels_one_gpu(precision):
    """b2n_slab_scatter / b2n_slab_gather (csrc/slab_exchange.cuh) with the "peer" grids of four
    slabs all on this GPU: the scatter stores every row of the local planes into every slab that
    holds it (halo rows into two slabs, wrap-around at the seam), the gather returns the sum over
    the slabs holding each row -- against the same bookkeeping done with torch slices."""
    import ctypes
    import torch
    from mrrt.nufft_b200 import _lib
    from mrrt.nufft_b200._slab import CudaSlabKernels, _pieces

    Nd, Kd, J = (20, 24, 18), (32, 40, 28), 6
    K1, K2, K3 = Kd
    k = CudaSlabKernels(Nd, Kd, (J, J, J), 1024, precision, False, (0.0, 0.0, 0.0), 1.0, None)
    cdt = torch.complex64 if precision == "single" else torch.complex128
    bounds = [0, 9, 17, 30, K2]
    slabs = [(bounds[s], bounds[s + 1] - bounds[s] + J - 1) for s in range(4)]     # (row0, rows + halo)
    grids = [torch.zeros((K3, nr, K1), dtype=cdt, device="cuda") for _, nr in slabs]
    ptrs = (ctypes.c_void_p * 4)(*[g.data_ptr() for g in grids])
    row0 = (ctypes.c_int * 4)(*[r for r, _ in slabs])
    nrows = (ctypes.c_int * 4)(*[n for _, n in slabs])
    z0, nz = 3, 7
    gen = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn((nz, K2, K1), dtype=cdt, device="cuda", generator=gen)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(k.lib.b2n_slab_scatter(k.gplan, ctypes.c_void_p(A.data_ptr()), nz, z0, 4, ptrs, row0, nrows, st))
    torch.cuda.synchronize()
    for (r0, nr), g in zip(slabs, grids):
        want = torch.zeros_like(g)
        for glo, llo, n in _pieces(r0, nr, K2):
            want[z0:z0 + nz, llo:llo + n] = A[:, glo:glo + n]
        assert torch.equal(g, want)
    # adjoint direction: every slab holds its own values; rows in two slabs add up
    for g in grids:
        g.copy_(torch.randn(g.shape, dtype=cdt, device="cuda", generator=gen))
    B = torch.full((nz, K2, K1), float("nan"), dtype=cdt, device="cuda")
    _lib.check(k.lib.b2n_slab_gather(k.gplan, ctypes.c_void_p(B.data_ptr()), nz, z0, 4, ptrs, row0, nrows, st))
    torch.cuda.synchronize()
    want = torch.zeros_like(B)
    for (r0, nr), g in zip(slabs, grids):
        for glo, llo, n in _pieces(r0, nr, K2):
            want[:, glo:glo + n] += g[z0:z0 + nz, llo:llo + n]
    tol = 1e-6 if precision == "single" else 1e-14
    assert float((B - want).abs().max() / want.abs().max()) <= tol
