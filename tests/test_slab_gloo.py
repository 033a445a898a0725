"""Host logic of the slab-distributed operator (mrrt/nufft_b200/_slab.py) over gloo on CPU,
world sizes 2 and 3: slab boundaries, sample partition, all-to-all packing, halo summation and
gathers.  The per-rank compute back end is the oracle (tests/slab_oracle.py); the result is
compared with the oracle's single-process NufftBase restatement."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from golden_util import rel_l2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _body(rank, q)
    except Exception:                       # report instead of leaving the parent waiting
        import traceback

        q.put({"rank": rank, "error": traceback.format_exc()})
    finally:
        dist.destroy_process_group()


def _body(rank, q):
    if True:
        from oracle import nufft_oracle as orc
        from mrrt.nufft_b200 import SlabShardedNufft
        from slab_oracle import OracleSlabKernels

        rs = np.random.RandomState(3)
        Nd, Kd, Jd = (10, 9, 7), (16, 14, 12), (4, 5, 4)
        n_shift = (1.0, 0.0, 2.5)
        M = 900
        om = (rs.rand(M, 3) * 2 - 1) * np.pi
        om[:40, 1] = np.pi - 1e-3          # windows that wrap around the periodic boundary
        om[40:80, 1] = -np.pi + 1e-3
        x = rs.standard_normal(Nd) + 1j * rs.standard_normal(Nd)
        y = rs.standard_normal(M) + 1j * rs.standard_normal(M)
        full = orc.OracleNufft(Nd=Nd, omega=om, Jd=Jd, Kd=Kd, precision="double", n_shift=n_shift)
        S = SlabShardedNufft(Nd, om, Jd=Jd, Kd=Kd, precision="double", n_shift=n_shift,
                             kernels=OracleSlabKernels(Nd, Kd, Jd, n_shift=n_shift), row_cost=5.0)
        res = {"rank": rank, "bounds": S.bounds, "M": S.M, "planes": (S.z0, S.z1)}
        yl = S.fft(x)
        res["fwd"] = rel_l2(yl, full.fft(x)[S.index])
        res["fwd_planes"] = rel_l2(S.fft(np.ascontiguousarray(x[:, :, S.z0:S.z1]), planes=True), yl)
        res["gather"] = rel_l2(S.gather_samples(yl), full.fft(x))
        xa = full.adj(y)
        res["adj"] = rel_l2(S.adj(y[S.index]), xa)
        res["adj_planes"] = rel_l2(S.adj(y[S.index], planes=True), xa[:, :, S.z0:S.z1])
        res["index_sum"] = int(S.index.sum())
        q.put(res)


@pytest.mark.parametrize("world", [2, 3])
def test_slab_sharding_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in out:
        assert "error" not in r, r["error"]
    assert all(p.exitcode == 0 for p in procs)
    out.sort(key=lambda r: r["rank"])
    assert all(r["bounds"] == out[0]["bounds"] for r in out)
    b = out[0]["bounds"]
    assert b[0] == 0 and b[-1] == 14 and all(b[i] < b[i + 1] for i in range(world))
    assert sum(r["M"] for r in out) == 900                       # a partition of the samples
    assert sum(r["index_sum"] for r in out) == 900 * 899 // 2
    assert out[0]["planes"][0] == 0 and out[-1]["planes"][1] == 7
    for r in out:
        for key in ("fwd", "fwd_planes", "gather", "adj", "adj_planes"):
            assert r[key] < 1e-11, (key, r)


def test_slab_boundaries_balance():
    import torch
    from mrrt.nufft_b200._slab import _pieces, row_statistics, slab_boundaries, window_rows

    rs = np.random.RandomState(0)
    Kd, Jd = (24, 384, 20), (6, 6, 6)
    om = np.clip(0.6 * rs.standard_normal((200000, 3)), -np.pi, np.pi - 1e-6).astype(np.float32)   # centre-heavy
    rows, n, cells = row_statistics(om, Jd, Kd, np.dtype(np.float32), torch.device("cpu"))
    assert np.array_equal(rows, window_rows(om[:, 1], 6, 384, np.dtype(np.float32)))
    assert rows.min() >= 0 and rows.max() < 384 and n.sum() == 200000
    # distinct origin cells per row, the slow way
    tm = om / np.array([2 * np.pi / k for k in Kd], dtype=np.float32)
    kw = np.mod(1 + np.floor(tm.astype(np.float64) - 3.0), Kd).astype(np.int64)
    uniq = np.unique(kw, axis=0)
    assert np.array_equal(cells, np.bincount(uniq[:, 1], minlength=384))
    cost = n + 1.38 * cells + 50.0
    for world in (2, 4, 8):
        b = slab_boundaries(cost, world)
        assert b[0] == 0 and b[-1] == 384 and all(b[i] < b[i + 1] for i in range(world))
        per = [cost[b[i]:b[i + 1]].sum() for i in range(world)]
        assert max(per) <= 1.15 * (sum(per) / world)
    assert _pieces(380, 10, 384) == [(380, 0, 4), (0, 4, 6)]
    assert _pieces(10, 5, 384) == [(10, 0, 5)]
