"""Host-side sharding logic over gloo, world_size 2, on CPU.  The per-rank operator is
the CPU oracle (tests may use it); the product's sharding classes are what is tested."""
import os
import socket

import numpy as np
import pytest
import torch
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
        from oracle import nufft_oracle as orc
        from mrrt.nufft_b200 import CoilShardedNufft, SampleShardedNufft

        rs = np.random.RandomState(0)
        Nd, Kd = (12, 10), (24, 20)
        om = (rs.rand(301, 2) * 2 - 1) * np.pi
        x = rs.standard_normal(Nd) + 1j * rs.standard_normal(Nd)
        y = rs.standard_normal(301) + 1j * rs.standard_normal(301)
        kw = dict(Jd=4, Kd=Kd, precision="double")
        full = orc.OracleNufft(Nd=Nd, omega=om, **kw)
        S = SampleShardedNufft(Nd, om, op_factory=orc.OracleNufft, **kw)
        res = {"rank": rank, "lo": S.lo, "hi": S.hi}
        res["fwd"] = rel_l2(S.fft(x), full.fft(x)[S.lo:S.hi])
        res["adj"] = rel_l2(S.adj(y[S.lo:S.hi]), full.adj(y))
        # sparse mode shards by matrix rows (= samples) with the same reduce
        fs = orc.OracleNufft(Nd=Nd, omega=om, mode="sparse", **kw)
        Ss = SampleShardedNufft(Nd, om, op_factory=orc.OracleNufft, mode="sparse", **kw)
        res["sp_fwd"] = rel_l2(Ss.fft(x), fs.fft(x)[Ss.lo:Ss.hi])
        res["sp_adj"] = rel_l2(Ss.adj(y[Ss.lo:Ss.hi]), fs.adj(y))
        res["sp_norm"] = rel_l2(Ss.norm(x), fs.adj(fs.fft(x)))
        xc = rs.standard_normal(Nd + (5,)) + 1j * rs.standard_normal(Nd + (5,))
        Cc = CoilShardedNufft(Nd, om, n_coils=5, op_factory=orc.OracleNufft, **kw)
        loc = Cc.fft(Cc.local_coils(xc))
        ref = full.fft(xc)[:, Cc.c0:Cc.c1]
        res["coil"] = rel_l2(loc.reshape(ref.shape), ref)
        res["coils"] = (Cc.c0, Cc.c1)
        q.put(res)
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions():
    from mrrt.nufft_b200 import shard_range

    for n in (0, 1, 7, 52707328):
        for w in (1, 2, 3, 8):
            edges = [shard_range(n, w, r) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_sample_and_coil_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out.sort(key=lambda r: r["rank"])
    assert (out[0]["lo"], out[0]["hi"], out[1]["lo"], out[1]["hi"]) == (0, 151, 151, 301)
    assert out[0]["coils"] == (0, 3) and out[1]["coils"] == (3, 5)
    for r in out:
        assert r["fwd"] < 1e-13 and r["adj"] < 1e-12 and r["coil"] < 1e-13
        assert r["sp_fwd"] < 1e-13 and r["sp_adj"] < 1e-12 and r["sp_norm"] < 1e-12
