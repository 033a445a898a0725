#!/usr/bin/env python
"""Benchmark of the NUFFT hot path (contract: see the task brief / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[4], the one `metric` is quoted on): 3-D 256^3 image,
3-D radial trajectory of 102 944 spokes x 512 samples (M = 52 707 328), Kd = 384^3,
Jd = 6 Kaiser-Bessel table (L = 1024), complex64, one coil.  A step is one forward
(`fft`) plus one adjoint (`adj`) transform; the metric is non-uniform points per second
= 2 * M / (t_fwd + t_adj).  Synthetic data (seeded normal), trajectory cast to float32
before the operator is built.

N > 1 (launched by torch.distributed.run, one rank per GPU).  Default `--sharding coils`:
an N-coil acquisition of the same volume, one coil per GPU, plan replicated, NO data-path
collective (weak scaling; value counts point-coils per second over all ranks).
`--sharding samples`: one coil, the spokes sharded over the ranks and the adjoint images
combined with one NCCL all-reduce per step (strong scaling; total work fixed).

`--impl reference` times the reference's own CPU implementation of the same path (its C
interpolators compiled unmodified into oracle/_ref, driven by the NumPy restatement of
its Python pipeline) on a bounded spoke subsample, on the host cores.
"""
import argparse
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
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 8 or f[0] != str(self.index):
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- CPU arm
class CpuReference(object):
    """The reference CPU path on a 1/frac spoke subsample (evenly strided spokes, so the
    sampling density pattern is preserved).  `step()` = one fft + one adj."""

    def __init__(self, frac=128):
        from oracle import nufft_oracle as orc

        self.engine = "reference" if orc.have_reference_engine() else "port"
        self.frac = frac
        # all the host threads this process may use, whatever OMP_NUM_THREADS says (torchrun
        # sets it to 1 for every rank): the reference's forward interpolator is OpenMP-parallel
        self.cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        try:
            import ctypes

            ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(self.cores))
        except OSError:  # pragma: no cover
            pass
        idx = np.arange(0, SPOKES, frac)
        om = np.concatenate([radial3d(SPOKES, NREAD, int(s), int(s) + 1) for s in idx], 0)
        t0 = time.perf_counter()
        self.O = orc.OracleNufft(Nd=ND, omega=om, Jd=JD, Kd=KD, precision="single",
                                 mode="table", engine=self.engine)
        self.t_plan = time.perf_counter() - t0
        self.x = image()
        self.Ms = om.shape[0]
        self.y = None
        self.t_if = self.t_ia = None

    def step(self):
        t0 = time.perf_counter(); self.y = self.O.fft(self.x); t_fwd = time.perf_counter() - t0
        t0 = time.perf_counter(); self.O.adj(self.y); t_adj = time.perf_counter() - t0
        return t_fwd, t_adj

    def interp_only(self):
        """Interpolation stages alone (grid_only), to extrapolate to the full M."""
        rs = np.random.RandomState(1)
        g = (rs.standard_normal(int(np.prod(KD))).astype(np.float32) + 0j).astype(np.complex64)
        t0 = time.perf_counter(); self.O.fft(g, grid_only=True); self.t_if = time.perf_counter() - t0
        t0 = time.perf_counter(); self.O.adj(self.y, grid_only=True); self.t_ia = time.perf_counter() - t0

    def summary(self, t_fwd, t_adj):
        M = SPOKES * NREAD
        fixed = max(t_fwd - self.t_if, 0.0) + max(t_adj - self.t_ia, 0.0)
        t_full = fixed + (self.t_if + self.t_ia) * (M / self.Ms)
        return {
            "kind": "reference" if self.engine == "reference" else "port",
            "pts_per_s_sample": 2 * self.Ms / (t_fwd + t_adj),
            "pts_per_s_full": 2 * M / t_full,
            "cores": self.cores,
            "sample": ("1/%d of the spokes (evenly strided, M=%d): fwd %.2fs (interp %.2fs, OpenMP over "
                       "samples) adj %.2fs (interp %.2fs, one thread per coil as in the reference), "
                       "numpy.fft for the 384^3 FFT; value = 2*M_full/(FFT+scaling time + interp time"
                       " * M_full/M_sample); on the sample itself %.3g points/s"
                       % (self.frac, self.Ms, t_fwd, self.t_if, t_adj, self.t_ia,
                          2 * self.Ms / (t_fwd + t_adj))),
        }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.perf_counter()
    ref = CpuReference(frac=args.cpu_frac)
    times = []
    for i in range(args.warmup + args.steps):
        tf, ta = ref.step()
        if i >= args.warmup:
            times.append((tf, ta))
    ref.interp_only()
    tf = float(np.mean([t[0] for t in times]))
    ta = float(np.mean([t[1] for t in times]))
    r = ref.summary(tf, ta)
    value = r["pts_per_s_full"]
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000 * (tf + ta),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "3D 256^3, 3-D radial 102944x512 (M=52707328), Kd=384^3, Jd=6, "
                               "table mode L=1024, complex64, 1 coil per GPU, fwd+adj per step (the "
                               "CPU transforms the coils one after another, so its points/s does "
                               "not depend on --gpus); each CPU step "
                               "runs a 1/%d spoke subsample" % args.cpu_frac},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------- GPU arm
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
    from mrrt.nufft_b200 import NufftBase, SampleShardedNufft, shard_range

    M = SPOKES * NREAD
    by_coil = world > 1 and args.sharding == "coils"
    if by_coil:
        s_lo, s_hi = 0, SPOKES                          # every rank: one whole coil of the volume
    else:
        s_lo, s_hi = shard_range(SPOKES, world, rank)   # shard whole spokes
    om_local = radial3d(SPOKES, NREAD, s_lo, s_hi)
    t0 = time.perf_counter()
    A = NufftBase(Nd=ND, omega=om_local, Jd=JD, Kd=KD, precision="single", mode="table",
                  on_gpu=True, device=dev, host_chunks=args.host_chunks)
    torch.cuda.synchronize()
    t_plan = time.perf_counter() - t0
    M_local = A.M

    def all_reduce_img(x):
        if world > 1 and not by_coil:
            mem = x.permute(2, 1, 0)
            dist.all_reduce(torch.view_as_real(mem), op=dist.ReduceOp.SUM)
        return x

    x_np = image(seed=rank if by_coil else 0)
    x_dev = torch.from_numpy(x_np).to(dev)              # F-ordered: consumed without a copy

    def step_dev():
        y = A.fft(x_dev)
        xa = A.adj(y)
        return all_reduce_img(xa)

    for _ in range(args.warmup):
        step_dev()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    A.set_option("profile", 1)
    A.kernel_timing()
    launches0 = A.launch_count
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_dev()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    kt = A.kernel_timing()
    A.set_option("profile", 0)
    launches = A.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    n_units = world if by_coil else 1                   # coils transformed per step, whole job
    value = 2.0 * M * n_units / (ms_step / 1000.0)

    # ---- end to end through the public API with pinned HOST buffers
    x_host = torch.from_numpy(x_np).pin_memory()
    x_host = x_host if x_host.stride() == torch.from_numpy(x_np).stride() else \
        torch.empty_strided(x_np.shape, torch.from_numpy(x_np).stride(), dtype=torch.complex64,
                            pin_memory=True).copy_(torch.from_numpy(x_np))

    def step_e2e_blocking():
        y_h = A.fft(x_host)                 # H2D image, transform, D2H samples
        if world == 1 or by_coil:
            return A.adj(y_h)               # H2D samples, transform, D2H image
        xa_d = all_reduce_img(A.adj(y_h.to(dev, non_blocking=True)))   # H2D samples, reduce
        out = torch.empty_strided(xa_d.shape, xa_d.stride(), dtype=xa_d.dtype, pin_memory=True)
        out.copy_(xa_d)                     # D2H image
        return out

    overlapped = (world == 1 or by_coil) and args.host_chunks > 1
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
        t = (time.perf_counter() - t0) / n
        tt = torch.tensor([t], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    n_e2e = max(2, min(args.steps, 10))
    t_e2e_blocking = time_e2e(step_e2e_blocking, n_e2e)
    t_e2e = time_e2e(step_e2e, n_e2e) if overlapped else t_e2e_blocking
    img_bytes = int(np.prod(ND)) * 8
    smp_bytes = M_local * 8
    e2e = {"value": 2.0 * M * n_units / t_e2e, "unit": UNIT, "ms_per_step": 1000 * t_e2e,
           "h2d_bytes_per_step": img_bytes + smp_bytes, "d2h_bytes_per_step": smp_bytes + img_bytes,
           "steps": n_e2e,
           "mode": ("fft(x_host, non_blocking=True); adj(k_host, non_blocking=True); synchronize() "
                    "-- pinned host buffers in and out, one synchronize per step, the two calls' "
                    "copies overlap in opposite link directions" if overlapped else
                    "blocking calls"),
           "blocking_ms_per_step": 1000 * t_e2e_blocking}

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
    PK = int(np.prod(KD))
    fwd_ms = kt[0] / max(kt[1], 1)
    adj_ms = kt[2] / max(kt[3], 1)
    # algorithmic bytes per launch: samples in (coords + value) and the
    # grid written once / read once
    adj_bytes = M_local * (3 * 4 + 8) + PK * 8
    fwd_bytes = PK * 8 + M_local * (3 * 4 + 8)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("adj_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    adj_gbs = adj_bytes / (adj_ms * 1e-3) / 1e9 if adj_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "spread_window3d_kernel<float,6> (adjoint gridding)",
                "achieved": adj_gbs, "peak": peak, "unit": "GB/s", "frac": adj_gbs / peak,
                "traffic": traffic, "peak_source": peak_src,
                "kernel_ms": adj_ms, "algorithmic_bytes": adj_bytes,
                "forward_kernel": {"kernel": "interp_fwd_tiled_kernel<float,3,6>", "kernel_ms": fwd_ms,
                                   "achieved": fwd_bytes / (fwd_ms * 1e-3) / 1e9 if fwd_ms > 0 else 0.0,
                                   "algorithmic_bytes": fwd_bytes},
                "whole_step": {"algorithmic_bytes": 6.00e9 if world == 1 else None,
                               "achieved": 6.00e9 / (ms_step * 1e-3) / 1e9 if world == 1 else None,
                               "frac": 6.00e9 / (ms_step * 1e-3) / 1e9 / peak if world == 1 else None},
                "note": "on-chip bound, not HBM bound: see DESIGN.md section 5"}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak" if (by_coil or world == 1) else "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "3D 256^3, 3-D radial 102944x512 (M=52707328), Kd=384^3, Jd=6, "
                               "table mode L=1024, complex64, 1 coil per GPU, fwd+adj per step",
                   "sharding": "none" if world == 1 else (
                       "coils: %d-coil acquisition, one coil per rank, no data-path collective; "
                       "value counts point-coils" % world if by_coil else
                       "samples sharded over %d ranks, NCCL all-reduce of the adjoint image" % world),
                   "l2": "inputs exceed L2 (grid 453 MB, samples 422 MB > 126 MB L2); no flush needed",
                   "plan_s": t_plan},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
    }
    if world == 1 and not args.no_cpu:
        try:
            ref = CpuReference(frac=args.cpu_frac)
            tf, ta = ref.step()
            ref.interp_only()
            r = ref.summary(tf, ta)
            out["cpu_baseline"] = {"value": r["pts_per_s_full"], "unit": UNIT, "cores": r["cores"],
                                   "kind": r["kind"], "sample": r["sample"]}
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
                    help="CPU legs use 1/frac of the spokes")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="c5", choices=["c5", "coils"],
                    help="c5 = BASELINE configs[4] (the contract bench, default); coils = "
                         "configs[3], the 32-coil 2-D batch sharded by coil")
    ap.add_argument("--sharding", default="coils", choices=["coils", "samples"],
                    help="N>1, c5 workload: coils = one coil of the volume per GPU, no collective "
                         "(weak scaling, default); samples = one coil, spokes sharded, NCCL "
                         "all-reduce of the adjoint image (strong scaling)")
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
