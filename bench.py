#!/usr/bin/env python
"""Benchmark of the NUFFT hot path (contract: see the task brief / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[4], the one `metric` is quoted on): 3-D 256^3 image,
3-D radial trajectory of 102 944 spokes x 512 samples (M = 52 707 328), Kd = 384^3,
Jd = 6 Kaiser-Bessel table (L = 1024), complex64, one coil.  A step is one forward
(`fft`) plus one adjoint (`adj`) transform; the metric is non-uniform points per second
= 2 * M / (t_fwd + t_adj).  Synthetic data (seeded normal), trajectory cast to float32
before the operator is built.

N > 1 (launched by torch.distributed.run, one rank per GPU) measures configs[4] itself: ONE
coil, the sample set sharded over the ranks (strong scaling, total work fixed).  Default
`--sharding slab` (SlabShardedNufft): image sharded by planes, oversampled grid and samples by
grid rows, one NCCL all-to-all per transform, halo rows summed by the receiver.
`--sharding samples` (SampleShardedNufft): spokes sharded, grid stage replicated, one NCCL
all-reduce of the adjoint image per step.  `--sharding coils`: an N-coil acquisition, one coil
per GPU, no collective (weak scaling); also reported as the secondary key `coil_replicas`.

`--impl reference` times the reference's own CPU implementation of the same path (its C
interpolators compiled unmodified into oracle/_ref, driven by the NumPy restatement of
its Python pipeline) on bounded spoke subsamples, on the host cores.
"""
import argparse
import datetime
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "non-uniform pts/s fwd+adj at 3D 256^3 Jd=6"
WORKLOAD = ("3D 256^3, 3-D radial 102944x512 (M=52707328), Kd=384^3, Jd=6, table mode L=1024, "
            "complex64, 1 coil, fwd+adj per step")
UNIT = "points/s"
ND = (256, 256, 256)
KD = (384, 384, 384)
JD = 6
SPOKES, NREAD = 102944, 512


def radial3d(spokes, nread, lo=0, hi=None, dtype=np.float32):
    """3-D radial trajectory, spoke directions on the golden-spiral sphere
    (SURVEY.md section 8d): z_s = 1-(2s+1)/S, phi_s = s*pi*(3-sqrt 5); samples
    r_i = 2*pi*(i-n/2)/n along each spoke.  Returns spokes [lo, hi)."""
    hi = spokes if hi is None else hi
    s = np.arange(lo, hi, dtype=np.float64)
    z = 1 - (2 * s + 1) / spokes
    phi = s * np.pi * (3 - np.sqrt(5))
    rxy = np.sqrt(1 - z * z)
    d = np.stack([rxy * np.cos(phi), rxy * np.sin(phi), z], 1)
    r = 2 * np.pi * (np.arange(nread) - nread // 2) / nread
    om = (d[:, None, :] * r[None, :, None]).reshape(-1, 3)
    return om.astype(dtype)


def image(seed=0):
    rs = np.random.RandomState(seed)
    x = rs.standard_normal(ND).astype(np.float32) + 1j * rs.standard_normal(ND).astype(np.float32)
    return np.asfortranarray(x.astype(np.complex64))


# --------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.

    nvidia-smi needs a few hundred ms to start, more than a short timed region lasts, so it is
    started before the warm-up steps and `wait_ready` blocks until its first line has arrived;
    every line carries nvidia-smi's own timestamp and only the samples taken between `mark_begin`
    and `mark_end` (host wall clock around the timed region) are reported.  If no sample falls
    inside a very short region, the samples of the 150 ms before / after it are used and
    `"window"` says so."""

    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None
        self.t_begin = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        try:
            for ln in self.proc.stdout:
                self.lines.append(ln)
        except Exception:  # pragma: no cover
            pass

    def wait_ready(self, timeout=3.0):
        t0 = time.perf_counter()
        while self.proc is not None and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def mark_begin(self):
        self.t_begin = datetime.datetime.now()

    def mark_end(self):
        self.t_end = datetime.datetime.now()

    @staticmethod
    def _stamp(text):
        try:
            return datetime.datetime.strptime(text, "%Y/%m/%d %H:%M:%S.%f")
        except ValueError:
            return None

    def stop(self):
        try:
            return self._stop()
        except Exception as e:  # the clocks line must never cost the bench line
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0,
                    "reasons": ["clock sampling failed: %s" % type(e).__name__]}

    def _stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for ln in list(self.lines):
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 9 or f[1] != str(self.index):
                continue
            try:
                rows.append((self._stamp(f[0]), float(f[2]), float(f[3]),
                             [nm for nm, val in zip(names, f[5:9]) if val.lower().startswith("active")]))
            except ValueError:
                continue
        window = "timed region"
        pick = rows
        if self.t_begin is not None and self.t_end is not None and all(r[0] is not None for r in rows):
            pick = [r for r in rows if self.t_begin <= r[0] <= self.t_end]
            if not pick:
                pad = datetime.timedelta(milliseconds=150)
                pick = [r for r in rows if self.t_begin - pad <= r[0] <= self.t_end + pad]
                window = "timed region +- 150 ms (no sample inside a region this short)"
        elif rows:
            window = "all samples since the warm-up (timestamps not usable)"
        sm = [r[1] for r in pick]
        mx = [r[2] for r in pick]
        reasons = set(nm for r in pick for nm in r[3])
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- CPU arm
def _set_omp_threads(n):
    try:
        import ctypes

        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except OSError:  # pragma: no cover
        pass


def spoke_subsample(frac):
    idx = np.arange(0, SPOKES, frac)
    return np.concatenate([radial3d(SPOKES, NREAD, int(s), int(s) + 1) for s in idx], 0)


class CpuReference(object):
    """The reference CPU path of the bench workload on bounded spoke subsamples (evenly strided
    spokes: the sampling density pattern is preserved).

    `step()` = one fft + one adj on 1/frac_step of the spokes (the 384^3 FFTs are done in
    full).  The interpolation stages alone are timed on the larger 1/frac_interp subsample
    (1/16) and their cost, linear in M, is scaled to the full M; the
    FFT + scaling part is the measured remainder of a step."""

    def __init__(self, frac_step=256, frac_interp=16):
        from oracle import nufft_oracle as orc

        self.orc = orc
        self.engine = "reference" if orc.have_reference_engine() else "port"
        self.frac_step, self.frac_interp = frac_step, frac_interp
        # all the host threads this process may use, whatever OMP_NUM_THREADS says (torchrun
        # sets it to 1 for every rank): the reference's forward interpolator is OpenMP-parallel
        self.cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        _set_omp_threads(self.cores)
        om = spoke_subsample(frac_step)
        self.O = orc.OracleNufft(Nd=ND, omega=om, Jd=JD, Kd=KD, precision="single",
                                 mode="table", engine=self.engine)
        self.x = image()
        self.Ms = om.shape[0]
        self.y = None

    def step(self):
        t0 = time.perf_counter(); self.y = self.O.fft(self.x); t_fwd = time.perf_counter() - t0
        t0 = time.perf_counter(); self.O.adj(self.y); t_adj = time.perf_counter() - t0
        return t_fwd, t_adj

    def _interp(self, O, threads):
        _set_omp_threads(threads)
        rs = np.random.RandomState(1)
        g = (rs.standard_normal(int(np.prod(KD))).astype(np.float32) + 0j).astype(np.complex64)
        ys = (rs.standard_normal(O.M).astype(np.float32) + 0j).astype(np.complex64)
        t0 = time.perf_counter(); O.fft(g, grid_only=True); t_f = time.perf_counter() - t0
        t0 = time.perf_counter(); O.adj(ys, grid_only=True); t_a = time.perf_counter() - t0
        _set_omp_threads(self.cores)
        return t_f, t_a

    def model(self, t_fwd, t_adj):
        """Full-workload step time from the measured pieces (see the class docstring)."""
        M = SPOKES * NREAD
        tf_s, ta_s = self._interp(self.O, self.cores)            # interpolation share of a step
        fixed = max(t_fwd - tf_s, 0.0) + max(t_adj - ta_s, 0.0)  # FFTs, scaling, phases
        Ob = self.orc.OracleNufft(Nd=ND, omega=spoke_subsample(self.frac_interp), Jd=JD, Kd=KD,
                                  precision="single", mode="table", engine=self.engine)
        tf_b, ta_b = self._interp(Ob, self.cores)
        tf_1, _ = self._interp(self.O, 1)                        # forward with ONE thread
        scale_b, scale_s = M / Ob.M, M / self.Ms
        t_full = fixed + (tf_b + ta_b) * scale_b
        t_full_1 = fixed + tf_1 * scale_s + ta_b * scale_b
        return {
            "kind": "reference" if self.engine == "reference" else "port",
            "cores": self.cores,
            "t_full_s": t_full,
            "value": 2 * M / t_full,
            "value_one_thread": 2 * M / t_full_1,
            "value_on_step_sample": 2 * self.Ms / (t_fwd + t_adj),
            "sample": ("step = fft+adj on 1/%d of the spokes (evenly strided, M=%d): fwd %.2fs adj %.2fs "
                       "incl. the full 384^3 numpy.fft; interpolation alone on 1/%d of the spokes "
                       "(M=%d): fwd %.2fs (OpenMP over samples, %d threads; %.2fs with 1 thread on the "
                       "step sample) adj %.2fs (one thread per coil, as the reference); value = "
                       "2*M_full / (FFT+scaling %.2fs + interpolation * M_full/M_sample)"
                       % (self.frac_step, self.Ms, t_fwd, t_adj, self.frac_interp, Ob.M, tf_b,
                          self.cores, tf_1, ta_b, fixed)),
        }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.perf_counter()
    ref = CpuReference(frac_step=args.cpu_frac, frac_interp=args.cpu_frac_interp)
    times = []
    for i in range(args.warmup + args.steps):
        tf, ta = ref.step()
        if i >= args.warmup:
            times.append((tf, ta))
    tf = float(np.mean([t[0] for t in times]))
    ta = float(np.mean([t[1] for t in times]))
    r = ref.model(tf, ta)
    # value and ms_per_step are the SAME quantity (full workload, modelled from the measured
    # pieces); what a timed step actually ran is reported next to them
    out = {
        "metric": METRIC, "value": r["value"], "unit": UNIT, "impl": "reference",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000 * r["t_full_s"], "extrapolated": True,
        "sample_ms_per_step": 1000 * (tf + ta), "value_on_step_sample": r["value_on_step_sample"],
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "sample": "each timed CPU step runs 1/%d of the spokes; see cpu_baseline.sample"
                             % args.cpu_frac},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"],
                         "one_thread": {"value": r["value_one_thread"], "cores": 1}},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------- GPU arm
def csrc_sha():
    """Hash of the kernel sources: ncu-derived numbers under profiles/ are only quoted for the
    sources they were measured on."""
    import glob
    import hashlib

    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(ROOT, "mrrt", "nufft_b200", "csrc", "*.cu*")) +
                    glob.glob(os.path.join(ROOT, "mrrt", "nufft_b200", "csrc", "*.h"))):
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def measured_traffic():
    """dram bytes per launch of the two interpolation kernels from the committed ncu capture
    (profiles/traffic.json, written by scripts/make_traffic.py), or None when the capture is of
    other kernel sources than the ones being benchmarked."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(tp))
    except Exception:
        return None, None, "no capture"
    if t.get("csrc_sha") != csrc_sha():
        return None, None, "capture is of other kernel sources (csrc_sha %s)" % t.get("csrc_sha")
    return (t.get("adj_kernel_dram_bytes_per_launch"), t.get("fwd_kernel_dram_bytes_per_launch"),
            t.get("source", "ncu --set full"))


def _pin(t):
    import torch

    out = torch.empty_strided(t.shape, t.stride(), dtype=t.dtype, pin_memory=True)
    out.copy_(t)
    return out


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from mrrt.nufft_b200 import NufftBase, SlabShardedNufft, shard_range

    M = SPOKES * NREAD
    sharding = "none" if world == 1 else args.sharding
    x_np = image(seed=rank if sharding == "coils" else 0)
    t0 = time.perf_counter()
    S = A = None
    if sharding == "slab":
        S = SlabShardedNufft(ND, radial3d(SPOKES, NREAD), Jd=JD, Kd=KD, precision="single", device=dev,
                             exchange=args.exchange)
        M_local = S.M
        slab_rows = S.nrows
        slab_exchange = S.exchange
        prof = S.k
    else:
        s_lo, s_hi = (0, SPOKES) if sharding in ("none", "coils") else shard_range(SPOKES, world, rank)
        A = NufftBase(Nd=ND, omega=radial3d(SPOKES, NREAD, s_lo, s_hi), Jd=JD, Kd=KD,
                      precision="single", mode="table", on_gpu=True, device=dev,
                      host_chunks=args.host_chunks)
        M_local = A.M
        prof = A
    torch.cuda.synchronize()
    t_plan = time.perf_counter() - t0
    try:        # device memory owned by the plan(s) of this rank, GB (before the e2e leg builds its sub-plans)
        plan_gb = (A.device_bytes if sharding != "slab" else
                   sum(int(S.k.lib.b2n_plan_device_bytes(q)) for q in (S.k.gplan, S.k.lplan) if q)) / 1e9
    except Exception:
        plan_gb = None

    # ---- device-resident step
    if sharding == "slab":
        xp_dev = torch.from_numpy(np.ascontiguousarray(x_np[:, :, S.z0:S.z1])).to(dev)

        def step_dev():
            return S.adj(S.fft(xp_dev, planes=True), planes=True)    # image stays sharded by planes
    else:
        x_dev = torch.from_numpy(x_np).to(dev)                       # F-ordered: consumed without a copy

        def step_dev():
            xa = A.adj(A.fft(x_dev))
            if sharding == "samples":
                dist.all_reduce(torch.view_as_real(xa.permute(2, 1, 0)), op=dist.ReduceOp.SUM)
            return xa

    sampler = ClockSampler(local_rank)
    if rank == 0:                       # nvidia-smi is up and sampling before the warm-up ends
        sampler.start()
        sampler.wait_ready()
    for _ in range(args.warmup):
        step_dev()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if S is not None:
        S.k.set_profile(1)
    else:
        A.set_option("profile", 1)
    prof.kernel_timing()
    launches0 = prof.launch_count
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    sampler.mark_begin()
    e0.record()
    for _ in range(args.steps):
        step_dev()
    e1.record()
    torch.cuda.synchronize()
    sampler.mark_end()
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    kt = prof.kernel_timing()
    if S is not None:
        S.k.set_profile(0)
    else:
        A.set_option("profile", 0)
    launches = prof.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    n_units = world if sharding == "coils" else 1                   # coils transformed per step, whole job
    value = 2.0 * M * n_units / (ms_step / 1000.0)

    # ---- end to end through the public API with pinned HOST buffers
    def time_e2e(fn, n):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        tt = torch.tensor([(time.perf_counter() - t0) / n], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    n_e2e = max(2, min(args.steps, 10))
    img_bytes = int(np.prod(ND)) * 8
    if sharding == "slab":
        xp_host = _pin(torch.from_numpy(np.ascontiguousarray(x_np[:, :, S.z0:S.z1])))
        k_host = _pin(S.fft(xp_dev, planes=True).cpu())
        y_out = torch.empty(S.M, dtype=torch.complex64, pin_memory=True)
        x_out = torch.empty_strided(xp_host.shape, xp_host.stride(), dtype=torch.complex64, pin_memory=True)

        def step_e2e_blocking():
            """This rank's image planes up, its samples down; its samples up, its planes down."""
            y = S.fft(xp_host.to(dev, non_blocking=True), planes=True)
            y_out.copy_(y, non_blocking=True)
            xa = S.adj(k_host.to(dev, non_blocking=True), planes=True)
            x_out.copy_(xa, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        s_h2d, s_d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

        def step_e2e():
            """The same four transfers on two copy streams (one per link direction) around the
            two transforms on the compute stream: the adjoint's samples travel up while the
            forward runs, the forward's samples travel down while the adjoint runs."""
            main = torch.cuda.current_stream()
            ev_x, ev_k, ev_y, ev_a = (torch.cuda.Event() for _ in range(4))
            with torch.cuda.stream(s_h2d):
                xd = xp_host.to(dev, non_blocking=True)
                ev_x.record(s_h2d)
                kd = k_host.to(dev, non_blocking=True)
                ev_k.record(s_h2d)
            xd.record_stream(main)
            kd.record_stream(main)
            main.wait_event(ev_x)
            y = S.fft(xd, planes=True)
            ev_y.record(main)
            y.record_stream(s_d2h)
            with torch.cuda.stream(s_d2h):
                s_d2h.wait_event(ev_y)
                y_out.copy_(y, non_blocking=True)
            main.wait_event(ev_k)
            xa = S.adj(kd, planes=True)
            ev_a.record(main)
            xa.record_stream(s_d2h)
            with torch.cuda.stream(s_d2h):
                s_d2h.wait_event(ev_a)
                x_out.copy_(xa, non_blocking=True)
            s_d2h.synchronize()
            main.synchronize()

        t_e2e_blocking = time_e2e(step_e2e_blocking, n_e2e)
        t_e2e = time_e2e(step_e2e, n_e2e)
        h2d = d2h = img_bytes * (S.z1 - S.z0) // ND[2] + S.M * 8
        mode = ("per rank: fft(planes) / adj(samples) from pinned host buffers, results back in pinned "
                "host buffers; host->device and device->host copies on their own streams around the "
                "transforms, one synchronize per step; bytes are this rank's (its image planes + its "
                "samples), every rank on its own PCIe link")
    else:
        x_host = _pin(torch.from_numpy(x_np))

        def step_e2e_blocking():
            y_h = A.fft(x_host)                 # H2D image, transform, D2H samples
            if sharding != "samples":
                return A.adj(y_h)               # H2D samples, transform, D2H image
            xa_d = A.adj(y_h.to(dev, non_blocking=True))
            dist.all_reduce(torch.view_as_real(xa_d.permute(2, 1, 0)), op=dist.ReduceOp.SUM)
            return _pin(xa_d)

        overlapped = sharding != "samples" and args.host_chunks > 1
        k_host = A.fft(x_host) if overlapped else None      # pinned samples: the adjoint's host input

        def step_e2e():
            """One fft and one adj from pinned HOST buffers; results back in pinned host memory.
            Both calls are issued non-blocking (like torch's .to(non_blocking=True)) and the step
            ends with ONE synchronize, so the samples of the fft travel device->host while the
            adjoint's samples travel host->device (the two link directions are independent)."""
            if not overlapped:
                return step_e2e_blocking()
            y_h = A.fft(x_host, non_blocking=True)
            xa_h = A.adj(k_host, non_blocking=True)
            A.synchronize()
            return y_h, xa_h

        t_e2e_blocking = time_e2e(step_e2e_blocking, n_e2e)
        t_e2e = time_e2e(step_e2e, n_e2e) if overlapped else t_e2e_blocking
        h2d = d2h = img_bytes + M_local * 8
        mode = ("fft(x_host, non_blocking=True); adj(k_host, non_blocking=True); synchronize() "
                "-- pinned host buffers in and out, one synchronize per step, the two calls' "
                "copies overlap in opposite link directions" if overlapped else "blocking calls")
    e2e = {"value": 2.0 * M * n_units / t_e2e, "unit": UNIT, "ms_per_step": 1000 * t_e2e,
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "h2d_gbs": h2d / t_e2e / 1e9, "d2h_gbs": d2h / t_e2e / 1e9,
           "steps": n_e2e, "mode": mode, "blocking_ms_per_step": 1000 * t_e2e_blocking}

    # ---- per-stage times of the slab operator (rank 0's, CUDA events between stages)
    stages = None
    if sharding == "slab":
        S.profile_stages(True)
        for _ in range(5):
            step_dev()
        mine = S.stage_times()
        S.profile_stages(False)
        mine["samples"] = int(S.M)
        every = [None] * world
        dist.all_gather_object(every, mine)
        stages = {"rank0": {k: v for k, v in mine.items() if k != "samples"},
                  "rows": [int(S.slabs[r][1]) for r in range(world)],
                  "samples": [e["samples"] for e in every],
                  "interp_ms_per_rank": [e["fwd_interp"] + e["adj_interp"] for e in every],
                  "axis3_ms_per_rank": [e["fwd_axis3"] + e["adj_axis3"] for e in every],
                  "all_to_all_ms_per_rank": [e["fwd_all_to_all"] + e["adj_all_to_all"] for e in every]}

    # ---- secondary (N > 1): the communication-free coil-replica number (weak scaling)
    secondary = None
    if world > 1 and sharding != "coils" and not args.no_secondary:
        S = A = prof = None
        torch.cuda.empty_cache()
        Ac = NufftBase(Nd=ND, omega=radial3d(SPOKES, NREAD), Jd=JD, Kd=KD, precision="single",
                       mode="table", on_gpu=True, device=dev)
        xc = torch.from_numpy(image(seed=rank)).to(dev)
        for _ in range(3):
            Ac.adj(Ac.fft(xc))
        torch.cuda.synchronize()
        dist.barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(5):
            Ac.adj(Ac.fft(xc))
        c1.record()
        torch.cuda.synchronize()
        tc = torch.tensor([c0.elapsed_time(c1) / 5], device=dev, dtype=torch.float64)
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        secondary = {"coil_replicas": {"ms_per_step": float(tc.item()), "scaling": "weak",
                                       "value": 2.0 * M * world / (float(tc.item()) / 1000.0),
                                       "note": "%d-coil acquisition, one coil per rank, plan replicated, "
                                               "no data-path collective; value counts point-coils" % world}}
        del Ac

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the adjoint gridding kernel)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    fwd_ms = kt[0] / max(kt[1], 1)
    adj_ms = kt[2] / max(kt[3], 1)
    # algorithmic bytes per launch, SURVEY 8(d): adjoint M(ndim r + c) + 2 P_K c (zero-fill and
    # final write; the timed interval covers the zero-fill), forward P_K c + M(ndim r + c); for a
    # slab P_K is this rank's slab
    grid_cells = int(np.prod(KD)) if sharding != "slab" else KD[0] * KD[2] * slab_rows
    adj_bytes = M_local * (3 * 4 + 8) + 2 * grid_cells * 8
    fwd_bytes = grid_cells * 8 + M_local * (3 * 4 + 8)
    tr_adj, tr_fwd, tr_src = measured_traffic() if world == 1 else (None, None, "single-GPU capture only")
    adj_gbs = adj_bytes / (adj_ms * 1e-3) / 1e9 if adj_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "spread_column3d_kernel<float,6> + zero-fill (adjoint gridding)",
                "achieved": adj_gbs, "peak": peak, "unit": "GB/s", "frac": adj_gbs / peak,
                "traffic": tr_adj, "traffic_source": tr_src, "peak_source": peak_src,
                "kernel_ms": adj_ms, "algorithmic_bytes": adj_bytes,
                "forward_kernel": {"kernel": "interp_fwd_tiled_kernel<float,3,6>", "kernel_ms": fwd_ms,
                                   "achieved": fwd_bytes / (fwd_ms * 1e-3) / 1e9 if fwd_ms > 0 else 0.0,
                                   "algorithmic_bytes": fwd_bytes, "traffic": tr_fwd},
                "whole_step": {"algorithmic_bytes": 6.00e9,
                               "achieved": 6.00e9 / (ms_step * 1e-3) / 1e9,
                               "frac": 6.00e9 / (ms_step * 1e-3) / 1e9 / (peak * world),
                               "note": "SURVEY 8(d) pair bytes / step time / (peak x n_gpus)"},
                "note": "on-chip bound, not HBM bound: see DESIGN.md section 5"}
    if sharding == "slab":
        roofline["note"] += "; kernel figures are rank 0's (its slab of the grid and its samples)"

    shard_txt = {
        "none": "none",
        "slab": "slab: image sharded by planes, grid + samples by grid rows over %d ranks, one "
                "exchange per transform (%s; SlabShardedNufft); image stays sharded between steps"
                % (world, "NCCL all-to-all" if sharding != "slab" or slab_exchange == "nccl" else
                   "direct stores/loads on peer memory over NVLink, device-side barriers"),
        "samples": "samples: spokes sharded over %d ranks, grid stage replicated, NCCL all-reduce of "
                   "the adjoint image" % world,
        "coils": "coils: %d-coil acquisition, one coil per rank, no data-path collective; value "
                 "counts point-coils" % world}[sharding]
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak" if sharding in ("none", "coils") else "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sharding": shard_txt,
                   "l2": "inputs exceed L2 (grid 453 MB, samples 422 MB > 126 MB L2); no flush needed",
                   "plan_s": t_plan, "plan_device_gb": plan_gb},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
    }
    if secondary:
        out["secondary"] = secondary
    if stages:
        out["slab_stages_ms"] = stages
    if world == 1 and not args.no_cpu:
        try:
            ref = CpuReference(frac_step=args.cpu_frac, frac_interp=args.cpu_frac_interp)
            tf, ta = ref.step()
            r = ref.model(tf, ta)
            out["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"],
                                   "kind": r["kind"], "sample": r["sample"], "extrapolated": True,
                                   "one_thread": {"value": r["value_one_thread"], "cores": 1}}
        except Exception as e:  # pragma: no cover
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(),
                                   "kind": "reference", "sample": "failed: %r" % (e,)}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_coil_arm(args):
    """BASELINE configs[3]: 2-D 320^2, 32-coil radial batch (503 spokes x 640 samples,
    Kd = 480^2, J = 6, complex64); coils sharded over the ranks, NO communication.
    Strong scaling of the 32-coil batch; value = 2 * M * 32 / step time."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from mrrt.nufft_b200 import CoilShardedNufft

    S, n, ncoil = 503, 640, 32
    ang = np.pi * np.arange(S) / S
    r = 2 * np.pi * (np.arange(n) - n / 2) / n
    om = np.stack([np.outer(np.cos(ang), r).ravel(), np.outer(np.sin(ang), r).ravel()], 1).astype(np.float32)
    A = CoilShardedNufft((320, 320), om, n_coils=ncoil, Jd=6, Kd=(480, 480), precision="single",
                         device=dev)
    nloc = A.c1 - A.c0
    g = torch.Generator(device=dev).manual_seed(rank)
    x = torch.randn((nloc, 320, 320), dtype=torch.complex64, device=dev, generator=g).permute(2, 1, 0)

    def step():
        return A.adj(A.fft(x))

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    if rank == 0:
        print(json.dumps({
            "metric": "non-uniform pts/s fwd+adj, 2D 320^2 32-coil batch", "unit": UNIT,
            "value": 2.0 * om.shape[0] * ncoil / (ms_step / 1000.0), "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "2D 320^2 radial 503x640 (M=321920), Kd=480^2, Jd=6, complex64, "
                                   "32 coils sharded over %d ranks (no communication)" % world}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-frac", type=int, default=256,
                    help="a timed CPU step runs 1/frac of the spokes")
    ap.add_argument("--cpu-frac-interp", type=int, default=16,
                    help="the CPU interpolation stages are timed on 1/frac of the spokes")
    ap.add_argument("--no-secondary", action="store_true",
                    help="N>1: skip the coil-replica secondary measurement")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="c5", choices=["c5", "coils"],
                    help="c5 = BASELINE configs[4] (the contract bench, default); coils = "
                         "configs[3], the 32-coil 2-D batch sharded by coil")
    ap.add_argument("--sharding", default="slab", choices=["slab", "samples", "coils"],
                    help="N>1, c5 workload: slab (default) = one coil, image by planes / grid and "
                         "samples by rows, NCCL all-to-all (strong scaling); samples = one coil, "
                         "spokes sharded, NCCL all-reduce of the adjoint image (strong scaling); "
                         "coils = one coil of the volume per GPU, no collective (weak scaling)")
    ap.add_argument("--exchange", default="auto", choices=["nccl", "p2p", "auto"],
                    help="slab sharding: how grid rows travel between the plane stage and the axis-3 "
                         "stage: NCCL all-to-all, or direct stores / loads on peer (symmetric) memory")
    ap.add_argument("--host-chunks", type=int, default=4,
                    help="sample ranges pipelined against host<->device copies in the e2e leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its
    # version banner to fd 1 when NCCL_DEBUG is set on the box), so fd 1 is pointed at
    # stderr for the duration of the run and the JSON line goes to the real stdout.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "coils":
        run_coil_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
