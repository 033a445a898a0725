"""Slab-distributed NUFFT of ONE large 3-D problem over the GPUs of a box (SURVEY 8(e) row 2
and 8(f)4): one process per GPU, ``torch.distributed`` (NCCL over NVLink / NVSwitch).

The reference is single-device (_nufft.py:1333-1369, :1526-1559); what is distributed here is
the path of ``NufftBase.fft`` / ``NufftBase.adj`` themselves:

* the IMAGE is sharded by planes of the last axis (rank r owns planes ``[z0_r, z1_r)``);
* the oversampled GRID is sharded by rows of axis 2: rank s owns the rows ``[b_s, b_s+1)`` of
  the window ORIGINS plus ``J-1`` halo rows, ``rows_s = (b_s + arange(width_s + J - 1)) mod K2``;
* the SAMPLES are sharded by the grid row of their window origin, so that every sample's whole
  J^3 window lies inside its rank's slab; slab boundaries balance a cost of samples + rows.

forward   x planes --(x*sn, zero-pad, batched 2-D FFT: b2n_planes_fwd)--> [K1, K2, nz_r]
          --(all-to-all: rows_s of every plane to rank s)--> [K1, rows_s, N3] zero-padded to K3
          --(FFT along axis 3 + phase_before: b2n_axis3_fwd)--> slab of the oversampled spectrum
          --(table interpolation on the slab plan: b2n_interp_fwd)--> this rank's samples
adjoint   the mirror image: b2n_interp_adj, b2n_axis3_adj, all-to-all (halo rows are summed by
          the receiver), b2n_planes_adj -> this rank's image planes (all-gathered on request).

Nothing is replicated except plan-time host work: each rank runs 1/G of every stage, and the
only communication is one all-to-all of ``N3*K1*K2*c / G`` bytes per rank and transform (the
sample-sharded operator replicates the whole scale + FFT + phase stage on every rank and
all-reduces the image).  Interpolation weights are bit-identical to the single-GPU plan's:
coordinates and window origins stay global, only grid rows are addressed locally (slab plan
of the C ABI, include/b200nufft.h).

``kernels`` is the per-rank compute back end: the CUDA library by default; the CPU tests pass
an oracle back end (tests/slab_oracle.py) to exercise this host logic over gloo.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from . import _plan_math as pm
from ._kernels import BeattyKernel
from ._sharded import shard_range

_TORCH_C = {np.dtype(np.complex64): torch.complex64, np.dtype(np.complex128): torch.complex128}


def _as3(v, typ):
    if np.isscalar(v):
        return (typ(v),) * 3
    if len(v) != 3:
        raise ValueError("SlabShardedNufft is 3-D: expected 3 values, got {}".format(len(v)))
    return tuple(typ(s) for s in v)


def window_rows(omega2, J, K, rdt):
    """Wrapped window origin along one axis for every sample, exactly as the device computes
    it: tm = omega / (2 pi / K) in the precision dtype (_nufft.py:338-342), origin
    1 + floor(tm - J/2.) in double (template.c:865-867), periodic wrap."""
    tm = np.asarray(omega2).astype(rdt, copy=False) / rdt.type(2 * np.pi / K)
    koff = 1 + np.floor(tm.astype(np.float64) - J / 2.0)
    return np.mod(koff, K).astype(np.int64)


def row_statistics(omega, Jd, Kd, rdt, device):
    """Per grid row of axis 2: ``rows`` (int64 [M], the row of every sample's window origin),
    ``n`` (samples per row) and ``cells`` (distinct window-origin cells per row).  Evaluated
    with torch on ``device`` (the GPU for the CUDA back end: a 52.7 M-sample trajectory takes
    ~0.1 s there instead of ~10 s of NumPy sorting on every rank)."""
    tdt = torch.float32 if rdt == np.float32 else torch.float64
    om = torch.from_numpy(np.ascontiguousarray(omega)).to(device=device, dtype=tdt)
    kw = []
    for d in range(3):
        # (T)(2 pi / K) as a one-element DEVICE tensor: with a 0-dim (scalar) divisor PyTorch
        # multiplies by the reciprocal instead of dividing, which differs from the IEEE division
        # the library (and the reference) performs by an ulp for some coordinates
        gam = torch.full((1,), 2 * np.pi / Kd[d], dtype=torch.float64).to(tdt).to(device)
        tm = torch.div(om[:, d], gam)
        koff = torch.floor(tm.to(torch.float64) - Jd[d] / 2.0).to(torch.int64) + 1
        kw.append(torch.remainder(koff, Kd[d]))
    rows = kw[1]
    K1, K2, K3 = Kd
    n = torch.bincount(rows, minlength=K2)
    key = (rows * K1 + kw[0]) * K3 + kw[2]
    key, _ = torch.sort(key)
    first = torch.ones_like(key, dtype=torch.bool)
    first[1:] = key[1:] != key[:-1]
    cells = torch.bincount(torch.div(key[first], K1 * K3, rounding_mode="floor"), minlength=K2)
    return rows.cpu().numpy(), n.cpu().numpy().astype(np.float64), cells.cpu().numpy().astype(np.float64)


# Cost of one origin row in units of "one sample in a dense region", fitted to the per-slab
# stage times of the 8-rank bench run (profiles/r02_slab_cost_fit.md, scripts/slab_cost_fit.py):
# the interpolation kernels cost 8.3e-8 ms per sample plus 5.1e-8 ms per occupied cell (a new
# cell is a forward slot that cannot be paired and, per group of columns, a window slide in the
# adjoint), the axis-3 passes 1.5e-3 ms per row of K1*K3 = 384^2 cells.
CELL_COST = 0.62
ROW_COST_PER_CELL = 18450.0 / (384.0 * 384.0)


def slab_boundaries(cost, world):
    """Boundaries ``b[0] = 0 < b[1] < ... < b[world] = K`` of the origin-row ranges that
    balance ``cost`` (one value per row) over the ranks; every range has at least one row."""
    cost = np.asarray(cost, dtype=np.float64)
    K = cost.shape[0]
    if world > K:
        raise ValueError("more ranks than grid rows")
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    b = [0]
    for s in range(1, world):
        target = cum[-1] * s / world
        k = int(np.searchsorted(cum, target, side="left"))
        # the boundary whose cumulative cost is closest to the target
        if k > 0 and abs(cum[k - 1] - target) < abs(cum[k] - target):
            k -= 1
        k = max(k, b[-1] + 1)
        k = min(k, K - (world - s))
        b.append(k)
    b.append(K)
    return b


def _pieces(row0, nrows, K):
    """The rows ``(row0 + arange(nrows)) mod K`` as at most two contiguous pieces:
    [(global_lo, local_lo, length), ...]."""
    first = min(nrows, K - row0)
    out = [(row0, 0, first)]
    if first < nrows:
        out.append((0, first, nrows - first))
    return out


class CudaSlabKernels(object):
    """Per-rank compute through the C ABI: a grid-stage plan with the GLOBAL geometry (no
    samples) for the plane stages and a SLAB plan (local rows, global coordinates) for the
    axis-3 stage and the interpolation."""

    def __init__(self, Nd, Kd, Jd, Ld, precision, ortho, n_shift, adjoint_scalefactor, device,
                 options=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA device required: mrrt.nufft_b200 has no CPU fallback")
        if device is None or (isinstance(device, torch.device) and device.index is None):
            device = torch.device("cuda", torch.cuda.current_device())
        elif not isinstance(device, torch.device):
            device = torch.device("cuda", int(device))
        self.device = device
        self.Nd, self.Kd, self.Jd, self.Ld = Nd, Kd, Jd, int(Ld)
        self.precision = precision
        self.rdt, self.cdt = pm.real_cplx_dtypes(precision)
        self.n_shift = n_shift
        self.options = dict(options or {})
        self.mids = pm.n_mid(Nd, "real")
        kernel = BeattyKernel(shape=Jd, grid_shape=Nd, os_grid_shape=Kd)
        self._sn1d = [np.ascontiguousarray(s, dtype=np.float64)
                      for s in pm.deapodization_1d(Nd, Kd, Jd, kernel.alpha, "real")]
        self._pb = [np.ascontiguousarray(a, dtype=self.rdt)
                    for a in pm.phase_before_angles(Kd, self.mids, self.rdt)]
        scale_ortho = float(np.sqrt(float(np.prod(Kd)))) if ortho else 1.0
        self.fwd_scale = 1.0 / scale_ortho if ortho else 1.0
        self.adj_scale = (adjoint_scalefactor / scale_ortho if ortho else float(adjoint_scalefactor))
        self._prec = _lib.B2N_SINGLE if precision == "single" else _lib.B2N_DOUBLE
        self.gplan = self._create(Nd, Kd, self._sn1d, self._pb)
        self.lplan = None
        self.M = 0

    def _create(self, Nd, Kd, sn1d, pb, opts=None):
        arr = lambda v: (ctypes.c_int * 3)(*v)
        plan = ctypes.c_void_p()
        _lib.check(self.lib.b2n_plan_create(3, arr(Nd), arr(Kd), arr(self.Jd), self.Ld, self._prec,
                                            0, self.device.index, ctypes.byref(plan)))
        for k, v in (opts or {}).items():
            _lib.check(self.lib.b2n_plan_set_option(plan, k.encode(), int(v)))
        sn_ptrs = (ctypes.c_void_p * 3)(*[s.ctypes.data for s in sn1d])
        pb_ptrs = (ctypes.c_void_p * 3)(*[a.ctypes.data for a in pb])
        _lib.check(self.lib.b2n_plan_set_scaling(plan, sn_ptrs, pb_ptrs, self.fwd_scale, self.adj_scale))
        return plan

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def make_local(self, omega_local, row0, nrows):
        """Slab plan over rows ``(row0 + arange(nrows)) mod K2`` and this rank's samples
        (``omega_local`` as handed in by the caller, any float dtype)."""
        K1, K2, K3 = self.Kd
        rows = (row0 + np.arange(nrows)) % K2
        # the slab plan never runs the image stages: unit deapodization; Nd[2] tells the fused
        # axis-3 pass how many planes are non-zero (forward) / survive the crop (adjoint)
        ones = [np.ones(1), np.ones(1), np.ones(self.Nd[2])]
        pb = [self._pb[0], np.ascontiguousarray(self._pb[1][rows]), self._pb[2]]
        opts = dict(self.options)
        if nrows != K2:
            opts.update({"slab_kglobal2": K2, "slab_origin2": row0})
        self.lplan = self._create((1, 1, self.Nd[2]), (K1, nrows, K3), ones, pb, opts)
        self.axis3_fused = bool(self.lib.b2n_plan_get_option(self.lplan, b"axis3_fused") == 1)
        self._keep = (ones, pb)
        om = np.asarray(omega_local)
        self.M = om.shape[0]
        om_rdt = np.ascontiguousarray(om.astype(self.rdt, copy=False).T)      # [3, M]
        hs = [np.ascontiguousarray(pm.lookup_table(self.Nd[d], self.Jd[d], self.Kd[d], self.Ld,
                                                   "real").astype(self.rdt)) for d in range(3)]
        ptrs = (ctypes.c_void_p * 3)(*[h.ctypes.data for h in hs])
        _lib.check(self.lib.b2n_plan_set_tables(self.lplan, ptrs))
        with torch.cuda.device(self.device):
            om_dev = torch.from_numpy(om_rdt).to(self.device)
            _lib.check(self.lib.b2n_plan_set_points(self.lplan, om_dev.data_ptr(), self.M,
                                                    _lib.B2N_COORD_OMEGA, self._stream()))
            # phase_after from omega as handed in (table mode: _nufft.py:313-315)
            ph = pm.phase_after(om, self.mids, self.n_shift, self.rdt, self.cdt)
            ph_dev = torch.from_numpy(np.ascontiguousarray(ph)).to(self.device)
            _lib.check(self.lib.b2n_plan_set_sample_phase(self.lplan, ph_dev.data_ptr(), self._stream()))
            torch.cuda.current_stream(self.device).synchronize()
        self.nrows = nrows

    def to_device(self, x):
        t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
        return t.to(self.device, dtype=_TORCH_C[self.cdt])

    def empty(self, shape):
        return torch.empty(shape, dtype=_TORCH_C[self.cdt], device=self.device)

    def planes_fwd(self, x_planes, z0):
        nz = x_planes.shape[0]
        out = self.empty((nz, self.Kd[1], self.Kd[0]))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b2n_planes_fwd(self.gplan, x_planes.data_ptr(), int(z0), int(nz),
                                               out.data_ptr(), self._stream()))
        return out

    def planes_adj(self, planes, z0):
        nz = planes.shape[0]
        out = self.empty((nz, self.Nd[1], self.Nd[0]))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b2n_planes_adj(self.gplan, planes.data_ptr(), int(z0), int(nz),
                                               out.data_ptr(), self._stream()))
        return out

    def axis3_fwd(self, grid):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b2n_axis3_fwd(self.lplan, grid.data_ptr(), self._stream()))

    def axis3_adj(self, grid):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b2n_axis3_adj(self.lplan, grid.data_ptr(), self._stream()))

    def interp_fwd(self, grid):
        out = self.empty((self.M,))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b2n_interp_fwd(self.lplan, grid.data_ptr(), out.data_ptr(), 1, 1,
                                               self._stream()))
        return out

    def interp_adj(self, samples, grid):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b2n_interp_adj(self.lplan, samples.data_ptr(), grid.data_ptr(), 1, 1,
                                               self._stream()))

    def option(self, name):
        return int(self.lib.b2n_plan_get_option(self.lplan, name.encode()))

    @property
    def launch_count(self):
        return int(self.lib.b2n_plan_launch_count(self.lplan)) + int(self.lib.b2n_plan_launch_count(self.gplan))

    def set_profile(self, on):
        _lib.check(self.lib.b2n_plan_set_option(self.lplan, b"profile", int(on)))

    def kernel_timing(self):
        out = (ctypes.c_double * 4)()
        _lib.check(self.lib.b2n_plan_get_timing(self.lplan, out))
        return tuple(out)

    def __del__(self):
        for name in ("lplan", "gplan"):
            plan = getattr(self, name, None)
            if plan is not None and plan.value:
                try:
                    self.lib.b2n_plan_destroy(plan)
                except Exception:  # pragma: no cover
                    pass
                setattr(self, name, None)


class SlabShardedNufft(object):
    """Slab-distributed 3-D NUFFT operator (table mode, real phasing).

    Every rank passes the SAME ``omega`` (the whole trajectory, ``(M, 3)``); rank r keeps the
    samples ``self.index`` (ascending global indices) whose window origins fall into its grid
    rows, and image planes ``[self.z0, self.z1)``.

    ``fft(x)``   ``x``: the whole image ``Nd`` (replicated; only this rank's planes are read) or,
                 with ``planes=True``, just this rank's planes ``Nd[:2] + (z1 - z0,)``.
                 Returns this rank's samples ``(len(index),)``.
    ``adj(k)``   ``k``: this rank's samples.  Returns the whole image on every rank (planes
                 all-gathered), or this rank's planes with ``planes=True``.
    Arrays: NumPy in -> NumPy out; torch tensors on the rank's device in -> tensors out."""

    def __init__(self, Nd, omega, Jd=4, Kd=None, precision="single", Ld=1024, ortho=False,
                 n_shift=None, adjoint_scalefactor=1.0, group=None, device=None, options=None,
                 kernels=None, row_cost=None, mode="table", phasing="real", on_gpu=True,
                 exchange="auto"):
        if mode != "table" or phasing != "real":
            raise ValueError("SlabShardedNufft supports mode='table' with phasing='real'")
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.Nd = _as3(Nd, int)
        self.Jd = _as3(Jd, int)
        self.Kd = _as3(Kd if Kd is not None else tuple(int(1.5 * n) for n in self.Nd), int)
        self.n_shift = _as3(n_shift if n_shift is not None else 0.0, float)
        omega = omega.cpu().numpy() if isinstance(omega, torch.Tensor) else np.asarray(omega)
        if omega.ndim != 2 or omega.shape[1] != 3:
            raise ValueError("number of cols must match NUFFT dimension")
        if omega.dtype not in (np.float32, np.float64):
            raise ValueError("omega must be float32 or float64")
        if precision == "auto":
            precision = "single" if omega.dtype == np.float32 else "double"
        self.precision = precision
        self.rdt, self.cdt = pm.real_cplx_dtypes(precision)
        self.M_total = omega.shape[0]
        N1, N2, N3 = self.Nd
        K1, K2, K3 = self.Kd
        J2 = self.Jd[1]
        G = self.world
        # ---- grid rows of the window origins -> per-row cost -> slab boundaries -> samples
        stat_dev = torch.device("cpu") if kernels is not None and getattr(kernels, "device", None) in (
            None, torch.device("cpu")) else (device if device is not None else torch.device(
                "cuda", torch.cuda.current_device()))
        if not isinstance(stat_dev, torch.device):
            stat_dev = torch.device("cuda", int(stat_dev))
        rows, n_row, cells_row = row_statistics(omega, self.Jd, self.Kd, self.rdt, stat_dev)
        if row_cost is None:
            row_cost = ROW_COST_PER_CELL * (K1 * K3)
        self.row_cost_model = n_row + CELL_COST * cells_row + float(row_cost)
        self.bounds = slab_boundaries(self.row_cost_model, G) if G > 1 else [0, K2]
        halo = J2 - 1 if G > 1 else 0
        self.slabs = []                                  # per rank: (row0, nrows)
        for s in range(G):
            w = self.bounds[s + 1] - self.bounds[s]
            if w + halo > K2:
                raise ValueError("grid too small for {} slabs with a halo of {} rows".format(G, halo))
            self.slabs.append((self.bounds[s], w + halo))
        b0, b1 = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.index = np.nonzero((rows >= b0) & (rows < b1))[0]
        self.M = int(self.index.shape[0])
        self.planes = [shard_range(N3, G, r) for r in range(G)]
        self.z0, self.z1 = self.planes[self.rank]
        self.row0, self.nrows = self.slabs[self.rank]
        if kernels is None:
            kernels = CudaSlabKernels(self.Nd, self.Kd, self.Jd, Ld, precision, ortho, self.n_shift,
                                      adjoint_scalefactor, device, options)
        self.k = kernels
        self.k.make_local(omega[self.index], self.row0, self.nrows)
        self.device = getattr(self.k, "device", torch.device("cpu"))
        self.exchange = "nccl"
        if exchange not in ("nccl", "p2p", "auto"):
            raise ValueError("exchange must be 'nccl', 'p2p' or 'auto'")
        if exchange != "nccl" and G > 1 and self.device.type == "cuda":
            self._setup_p2p(strict=exchange == "p2p")
        # all-to-all split sizes (complex elements)
        nz_me = self.z1 - self.z0
        self._fwd_in = [nz_me * self.slabs[s][1] * K1 for s in range(G)]
        self._fwd_out = [(self.planes[r][1] - self.planes[r][0]) * self.nrows * K1 for r in range(G)]

    # ------------------------------------------------------------------ stage timing
    def profile_stages(self, on=True):
        """Record CUDA events between the stages of every transform (CUDA back end only)."""
        self._stage_ev = [] if on and self.device.type == "cuda" else None

    def _mark(self, name):
        if getattr(self, "_stage_ev", None) is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(self.device))
            self._stage_ev.append((name, ev))

    def stage_times(self):
        """Mean milliseconds per stage since ``profile_stages()`` (synchronises)."""
        evs = getattr(self, "_stage_ev", None) or []
        torch.cuda.synchronize(self.device)
        tot, cnt = {}, {}
        for (n0, e0), (n1, e1) in zip(evs[:-1], evs[1:]):
            if n1 == "begin":
                continue
            tot[n1] = tot.get(n1, 0.0) + e0.elapsed_time(e1)
            cnt[n1] = cnt.get(n1, 0) + 1
        self._stage_ev = []
        return {k: tot[k] / cnt[k] for k in tot}

    # ------------------------------------------------------------------ peer-memory exchange
    def _setup_p2p(self, strict):
        """Slab grids in SYMMETRIC memory (torch.distributed._symmetric_memory: every rank's
        buffer is mapped into every other rank's address space over NVLink).  The exchange
        between the plane stage and the axis-3 stage then needs no pack pass, no NCCL staging
        and no unpack pass: in the forward transform each rank stores the rows of its planes
        straight into the destination ranks' grids; in the adjoint each rank reads its planes
        straight out of the source ranks' grids, adding the halo rows as it goes.  Device-side
        barriers (signal pads) order the phases."""
        ok = torch.ones(1, device=self.device)
        try:
            import torch.distributed._symmetric_memory as symm_mem

            K1, K2, K3 = self.Kd
            if (K1 * np.dtype(self.cdt).itemsize) % 16 != 0 or self.world > 16:
                raise RuntimeError("the exchange kernels move 16-byte vectors and know at most 16 peers")
            nmax = max(nrows for _, nrows in self.slabs)
            tdt = _TORCH_C[self.cdt]
            rdt = torch.float32 if tdt == torch.complex64 else torch.float64
            grp = self.group if self.group is not None else dist.group.WORLD
            self._sym = symm_mem.empty(K3 * nmax * K1 * 2, dtype=rdt, device=self.device)
            self._hdl = symm_mem.rendezvous(self._sym, grp)
            self._peer_grids = []
            for r in range(self.world):
                nr = self.slabs[r][1]
                buf = self._hdl.get_buffer(r, (K3, nr, K1, 2), rdt, 0)
                self._peer_grids.append(torch.view_as_complex(buf))
            # argument arrays of the exchange kernels (b2n_slab_scatter / b2n_slab_gather)
            self._pg_ptrs = (ctypes.c_void_p * self.world)(*[g.data_ptr() for g in self._peer_grids])
            self._pg_row0 = (ctypes.c_int * self.world)(*[int(r0) for r0, _ in self.slabs])
            self._pg_nrows = (ctypes.c_int * self.world)(*[int(nr) for _, nr in self.slabs])
        except Exception:
            if strict:
                raise
            ok.zero_()
        # every rank must take the same path
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if ok.item() > 0:
            self.exchange = "p2p"
        elif strict:
            raise RuntimeError("symmetric-memory exchange is not available on every rank")

    def _barrier(self):
        self._hdl.barrier(channel=0)

    def _fft_dev_p2p(self, xp):
        K1, K2, K3 = self.Kd
        N3 = self.Nd[2]
        self._mark("begin")
        A = self.k.planes_fwd(xp, self.z0)                        # [nz, K2, K1]
        self._mark("fwd_planes")
        self._barrier()                                           # every rank is done with its grid
        # ONE kernel: the rows of this rank's planes stored straight into every slab that holds
        # them (csrc/slab_exchange.cuh) -- all-to-all, pack and unpack in a single pass
        _lib.check(self.k.lib.b2n_slab_scatter(self.k.gplan, ctypes.c_void_p(A.data_ptr()), self.z1 - self.z0,
                                               self.z0, self.world, self._pg_ptrs, self._pg_row0,
                                               self._pg_nrows, self.k._stream()))
        grid = self._peer_grids[self.rank]
        if not getattr(self.k, "axis3_fused", False):
            grid[N3:].zero_()
        self._barrier()                                           # all rows have landed
        self._mark("fwd_all_to_all")
        self.k.axis3_fwd(grid)
        self._mark("fwd_axis3")
        y = self.k.interp_fwd(grid)
        self._mark("fwd_interp")
        return y

    def _adj_dev_p2p(self, kt):
        K1, K2, K3 = self.Kd
        self._mark("begin")
        grid = self._peer_grids[self.rank]
        self.k.interp_adj(kt, grid)
        self._mark("adj_interp")
        self.k.axis3_adj(grid)
        self._mark("adj_axis3")
        nz = self.z1 - self.z0
        B = self.k.empty((nz, K2, K1))
        self._barrier()                                           # every slab is gridded and transformed
        # ONE kernel: this rank's planes read out of every slab, the rows held by more than one
        # slab (halos) summed in slab order on the way
        _lib.check(self.k.lib.b2n_slab_gather(self.k.gplan, ctypes.c_void_p(B.data_ptr()), nz, self.z0,
                                              self.world, self._pg_ptrs, self._pg_row0, self._pg_nrows,
                                              self.k._stream()))
        self._barrier()                                           # the grids may be reused
        self._mark("adj_all_to_all")
        out = self.k.planes_adj(B, self.z0)
        self._mark("adj_planes")
        return out

    # ------------------------------------------------------------------ helpers
    def _a2a(self, out, inp, out_splits, in_splits):
        if self.world == 1:
            out.copy_(inp)
            return
        dist.all_to_all_single(torch.view_as_real(out), torch.view_as_real(inp), out_splits,
                               in_splits, group=self.group)

    def _image_planes(self, x, planes):
        """This rank's planes as a contiguous [nz, N2, N1] tensor on the compute device."""
        N1, N2, N3 = self.Nd
        nz = self.z1 - self.z0
        if isinstance(x, torch.Tensor):
            want = (N1, N2, nz) if planes else (N1, N2, N3)
            if tuple(x.shape) != want:
                raise ValueError("expected an image of shape {}".format(want))
            mem = x.permute(2, 1, 0)
            if not planes:
                mem = mem[self.z0:self.z1]
            return self.k.to_device(mem.contiguous())
        x = np.asarray(x)
        want = (N1, N2, nz) if planes else (N1, N2, N3)
        if x.shape != want:
            raise ValueError("expected an image of shape {}".format(want))
        mem = x.transpose(2, 1, 0)
        if not planes:
            mem = mem[self.z0:self.z1]
        return self.k.to_device(np.ascontiguousarray(mem))

    # ------------------------------------------------------------------ transforms
    def fft(self, x, planes=False):
        """Forward transform; returns this rank's samples (see the class docstring)."""
        is_np = not isinstance(x, torch.Tensor)
        xp = self._image_planes(x, planes)
        y = self._fft_dev(xp)
        return y.cpu().numpy() if is_np else y

    def _fft_dev(self, xp):
        if self.exchange == "p2p":
            return self._fft_dev_p2p(xp)
        K1, K2, K3 = self.Kd
        N3 = self.Nd[2]
        self._mark("begin")
        A = self.k.planes_fwd(xp, self.z0)                        # [nz, K2, K1]
        self._mark("fwd_planes")
        nz = A.shape[0]
        send = self.k.empty((sum(self._fwd_in),))
        off = 0
        for s in range(self.world):
            row0, nrows = self.slabs[s]
            view = send[off:off + nz * nrows * K1].view(nz, nrows, K1)
            for glo, llo, n in _pieces(row0, nrows, K2):
                view[:, llo:llo + n].copy_(A[:, glo:glo + n])
            off += nz * nrows * K1
        grid = self.k.empty((K3, self.nrows, K1))
        if not getattr(self.k, "axis3_fused", False):
            grid[N3:].zero_()             # (the fused axis-3 pass creates the padding itself)
        self._mark("fwd_pack")
        self._a2a(grid[:N3].view(-1), send, self._fwd_out, self._fwd_in)
        self._mark("fwd_all_to_all")
        self.k.axis3_fwd(grid)
        self._mark("fwd_axis3")
        y = self.k.interp_fwd(grid)
        self._mark("fwd_interp")
        return y

    def adj(self, k_local, planes=False):
        """Adjoint transform of this rank's samples (see the class docstring)."""
        is_np = not isinstance(k_local, torch.Tensor)
        kt = self.k.to_device(k_local).reshape(-1)
        if kt.numel() != self.M:
            raise ValueError("invalid size")
        xp = self._adj_dev(kt)                                    # [nz, N2, N1]
        if not planes:
            xp = self._gather_planes(xp)
        out = xp.permute(2, 1, 0)
        return out.cpu().numpy() if is_np else out

    def _adj_dev(self, kt):
        if self.exchange == "p2p":
            return self._adj_dev_p2p(kt)
        K1, K2, K3 = self.Kd
        N3 = self.Nd[2]
        self._mark("begin")
        grid = self.k.empty((K3, self.nrows, K1))
        self.k.interp_adj(kt, grid)
        self._mark("adj_interp")
        self.k.axis3_adj(grid)
        self._mark("adj_axis3")
        nz = self.z1 - self.z0
        recv = self.k.empty((sum(self._fwd_in),))
        self._a2a(recv, grid[:N3].view(-1), self._fwd_in, self._fwd_out)
        self._mark("adj_all_to_all")
        # every grid row is the ORIGIN row of exactly one slab: those parts are copied (no
        # zero-fill pass), then the halo rows of the neighbouring slabs are added
        B = self.k.empty((nz, K2, K1))
        views, off = [], 0
        for s in range(self.world):
            row0, nrows = self.slabs[s]
            views.append(recv[off:off + nz * nrows * K1].view(nz, nrows, K1))
            off += nz * nrows * K1
            own = self.bounds[s + 1] - self.bounds[s]
            B[:, row0:row0 + own].copy_(views[s][:, :own])
        for s in range(self.world):
            row0, nrows = self.slabs[s]
            own = self.bounds[s + 1] - self.bounds[s]
            if nrows > own:
                for glo, llo, n in _pieces((row0 + own) % K2, nrows - own, K2):
                    B[:, glo:glo + n] += views[s][:, own + llo:own + llo + n]
        self._mark("adj_unpack")
        out = self.k.planes_adj(B, self.z0)
        self._mark("adj_planes")
        return out

    def _all_gather_padded(self, t, counts):
        """All-gather of per-rank 1-D complex tensors of (possibly) different lengths: every
        rank's piece is padded to the longest one (NCCL all-gather wants equal sizes)."""
        mx = max(counts)
        pad = self.k.empty((mx,))
        pad[:t.numel()].copy_(t)
        buf = self.k.empty((self.world * mx,))
        dist.all_gather_into_tensor(torch.view_as_real(buf), torch.view_as_real(pad), group=self.group)
        return [buf[r * mx:r * mx + c] for r, c in enumerate(counts)]

    def _gather_planes(self, xp):
        if self.world == 1:
            return xp
        N1, N2, N3 = self.Nd
        full = self.k.empty((N3, N2, N1))
        sizes = [(b - a) * N2 * N1 for a, b in self.planes]
        if len(set(sizes)) == 1:
            dist.all_gather_into_tensor(torch.view_as_real(full.view(-1)),
                                        torch.view_as_real(xp.reshape(-1)), group=self.group)
        else:
            off = 0
            for piece, n in zip(self._all_gather_padded(xp.reshape(-1), sizes), sizes):
                full.view(-1)[off:off + n].copy_(piece)
                off += n
        return full

    def norm(self, x):
        return self.adj(self.fft(x))

    def gather_samples(self, k_local):
        """All ranks' samples assembled in acquisition order (test / convenience helper)."""
        is_np = not isinstance(k_local, torch.Tensor)
        kt = self.k.to_device(k_local).reshape(-1)
        out = self.k.empty((self.M_total,))
        if self.world == 1:
            out[torch.as_tensor(self.index, device=out.device)] = kt
        else:
            info = [None] * self.world
            dist.all_gather_object(info, self.index, group=self.group)
            pieces = self._all_gather_padded(kt, [len(i) for i in info])
            for idx, piece in zip(info, pieces):
                out[torch.as_tensor(idx, device=out.device)] = piece
        return out.cpu().numpy() if is_np else out
