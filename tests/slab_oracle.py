"""CPU back end for ``SlabShardedNufft`` built on the oracle -- TEST INFRASTRUCTURE ONLY.

Implements the ``kernels`` interface of mrrt/nufft_b200/_slab.py (planes_fwd, axis3_fwd,
interp_fwd, interp_adj, axis3_adj, planes_adj) with NumPy and the oracle's C interpolators, so
that the slab partition, the all-to-all packing, the halo summation and the gathers can run over
gloo on CPU.  Stage arithmetic restates the reference pipeline (_nufft.py:1325-1391, :1495-1572)
split at the same places as the C ABI's staged transforms (include/b200nufft.h)."""
import numpy as np
import torch

from oracle import nufft_oracle as orc


class OracleSlabKernels(object):
    def __init__(self, Nd, Kd, Jd, Ld=1024, precision="double", n_shift=(0.0, 0.0, 0.0), engine="port"):
        self.Nd, self.Kd, self.Jd, self.Ld = tuple(Nd), tuple(Kd), tuple(Jd), Ld
        self.rdt = np.dtype(np.float32 if precision == "single" else np.float64)
        self.cdt = np.dtype(np.complex64 if precision == "single" else np.complex128)
        self.n_shift = tuple(n_shift)
        self.engine = engine
        self.mids = orc.n_mid_of(self.Nd, "real")
        self.sn = orc.scaling_factors(self.Nd, self.Kd, self.Jd, "real").astype(self.rdt)   # [N1,N2,N3]
        self.pb = orc.phase_before(self.Kd, self.mids, self.rdt, self.cdt)                  # [K1,K2,K3]
        self.h = [np.real(orc.make_table(self.Nd[d], self.Jd[d], self.Kd[d], Ld, "real")).astype(self.rdt)
                  for d in range(3)]
        self.device = torch.device("cpu")

    # ---- interface
    def make_local(self, omega_local, row0, nrows):
        K1, K2, K3 = self.Kd
        self.row0, self.nrows = row0, nrows
        self.rows = (row0 + np.arange(nrows)) % K2
        om = np.asarray(omega_local)
        self.M = om.shape[0]
        self.phase_after = orc.phase_after(om, self.mids, self.n_shift, self.rdt, self.cdt)
        om = om.astype(self.rdt, copy=False)
        tm = np.zeros(om.shape, dtype=self.rdt, order="F")
        for d in range(3):
            tm[:, d] = om[:, d] / (2 * np.pi / self.Kd[d])
        # axis 2 in LOCAL row units: shift by (local origin row - unwrapped origin)
        koff = 1 + np.floor(tm[:, 1].astype(np.float64) - self.Jd[1] / 2.0)
        kloc = np.mod(np.mod(koff, K2) - row0, K2)
        assert self.M == 0 or (kloc + self.Jd[1] <= nrows).all(), "sample outside the slab"
        tm[:, 1] = tm[:, 1] + (kloc - koff)
        self.tm = tm
        self.Kloc = (K1, nrows, K3)

    def to_device(self, x):
        t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
        return t.to(torch.complex64 if self.cdt == np.complex64 else torch.complex128)

    def empty(self, shape):
        return torch.empty(shape, dtype=torch.complex64 if self.cdt == np.complex64 else torch.complex128)

    def planes_fwd(self, x_planes, z0):
        x = x_planes.numpy()                                       # [nz, N2, N1]
        nz = x.shape[0]
        sn = self.sn.transpose(2, 1, 0)[z0:z0 + nz]
        out = np.fft.fftn(x * sn, s=(self.Kd[1], self.Kd[0]), axes=(1, 2)).astype(self.cdt)
        return torch.from_numpy(np.ascontiguousarray(out))

    def planes_adj(self, planes, z0):
        p = planes.numpy()                                         # [nz, K2, K1]
        nz = p.shape[0]
        x = np.fft.ifftn(p, axes=(1, 2)) * (self.Kd[0] * self.Kd[1])
        x = x[:, :self.Nd[1], :self.Nd[0]] * np.conj(self.sn.transpose(2, 1, 0)[z0:z0 + nz])
        return torch.from_numpy(np.ascontiguousarray(x.astype(self.cdt)))

    def axis3_fwd(self, grid):
        g = grid.numpy()                                           # [K3, nrows, K1], in place
        g[...] = np.fft.fft(g, axis=0) * self.pb.transpose(2, 1, 0)[:, self.rows, :]

    def axis3_adj(self, grid):
        g = grid.numpy()
        g[...] = np.fft.ifft(g * np.conj(self.pb.transpose(2, 1, 0)[:, self.rows, :]), axis=0) * self.Kd[2]

    def interp_fwd(self, grid):
        g = grid.numpy().reshape(-1)                               # memory order = F-order of the slab
        y = orc.interp_table(self.Kloc, self.Jd, self.Ld, self.h, self.tm, g, engine=self.engine)[:, 0]
        return torch.from_numpy(np.ascontiguousarray((y * self.phase_after).astype(self.cdt)))

    def interp_adj(self, samples, grid):
        y = (samples.numpy() * np.conj(self.phase_after)).astype(self.cdt)
        g = orc.interp_table_adj(self.Kloc, self.Jd, self.Ld, self.h, self.tm, y, engine=self.engine)[:, 0]
        grid.numpy()[...] = g.reshape(grid.shape)
