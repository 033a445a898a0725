"""Shared-memory bank-conflict model of the tiled forward kernel on the bench trajectory.

Row pitch 32 complex64 == 0 mod 16 bank pairs, so for every tap the bank pair of a lane is
(x1 + j1) mod 16: the wavefronts of one LDS.64 of a half-warp = max over columns c of the
number of DISTINCT cells with x1 mod 16 == c among its 16 lanes.  Compares the cell-sorted
order with the column-interleaved order (rank within column slowest, column fastest), for
plain samples and for slots (pairs of same-cell samples, option fwd_pair)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import bench

TILE = (16, 8, 8)
K = 384
J = 6
om = bench.radial3d(bench.SPOKES, bench.NREAD)
tm = om / np.float32(2 * np.pi / K)
kw = (1 + np.floor(tm.astype(np.float64) - J / 2.0)).astype(np.int64) % K
b = [kw[:, d] // TILE[d] for d in range(3)]
c = [kw[:, d] % TILE[d] for d in range(3)]
nb = [K // t for t in TILE]
binid = b[0] + nb[0] * (b[1] + nb[1] * b[2])
cell = c[0] + TILE[0] * (c[1] + TILE[1] * c[2])
rs = np.random.RandomState(0)
sel_bins = rs.choice(nb[0] * nb[1] * nb[2], 400, replace=False)


def factor(cells_in_order):
    n = len(cells_in_order)
    tot, cnt = 0, 0
    for s in range(0, n, 16):
        h = cells_in_order[s:s + 16]
        u = np.unique(h)
        col = u % 16
        tot += np.bincount(col, minlength=16).max()
        cnt += 1
    return tot, cnt


def interleave(cs):
    col = cs % 16
    rank = np.zeros(len(cs), dtype=np.int64)
    o2 = np.argsort(col, kind="stable")
    sc = col[o2]
    start = np.searchsorted(sc, np.arange(16))
    rank[o2] = np.arange(len(cs)) - start[sc]
    return cs[np.lexsort((col, rank))]


res = {k: [0, 0] for k in ("samples sorted", "samples interleaved", "slots sorted", "slots interleaved")}
nsamp = 0
for bb in sel_bins:
    idx = np.nonzero(binid == bb)[0]
    if len(idx) == 0:
        continue
    cs = np.sort(cell[idx])
    u, cnt = np.unique(cs, return_counts=True)
    slots = np.repeat(u, (cnt + 1) // 2)
    for name, arr in (("samples sorted", cs), ("samples interleaved", interleave(cs)),
                      ("slots sorted", slots), ("slots interleaved", interleave(slots))):
        a, n = factor(arr)
        res[name][0] += a
        res[name][1] += n
    nsamp += len(idx)
print("samples", nsamp)
for k, (a, n) in res.items():
    print("%-20s half-warp LDS instructions %7d  wavefronts %7d  factor %.3f" % (k, n, a, a / n))
