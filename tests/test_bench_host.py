"""Host-side pieces of bench.py that can be checked without a GPU: the synthetic trajectory,
the clock sampler's parsing / windowing (against a fake nvidia-smi), and that the ncu-derived
traffic figures under profiles/ belong to the kernel sources in the tree."""
import json
import os
import stat
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_radial3d_ranges_are_slices_of_the_whole():
    """Spoke ranges (what the slab / sample shardings and the CPU subsamples build) are exactly
    the rows of the full trajectory; coordinates stay inside [-pi, pi)."""
    S, n = 257, 16
    full = bench.radial3d(S, n)
    assert full.shape == (S * n, 3) and full.dtype == np.float32
    assert np.array_equal(bench.radial3d(S, n, 40, 97), full[40 * n:97 * n])
    assert np.array_equal(bench.radial3d(S, n, 5, 6), full[5 * n:6 * n])
    assert float(full.min()) >= -np.pi and float(full.max()) < np.pi
    # every spoke passes through the origin at its middle sample and is a straight line
    mid = full.reshape(S, n, 3)[:, n // 2]
    assert np.all(mid == 0)
    r = np.linalg.norm(full.reshape(S, n, 3).astype(np.float64), axis=2)
    assert np.allclose(r[:, 0], np.pi, rtol=1e-6)


def test_bench_workload_constants():
    assert bench.SPOKES * bench.NREAD == 52707328
    assert bench.ND == (256, 256, 256) and bench.KD == (384, 384, 384) and bench.JD == 6
    assert "M=52707328" in bench.WORKLOAD


def _fake_smi(tmp_path, body):
    p = tmp_path / "nvidia-smi"
    p.write_text("#!%s\n%s" % (sys.executable, body))
    p.chmod(p.stat().st_mode | stat.S_IEXEC)
    return str(tmp_path)


_FAKE = '''
import datetime, time
while True:
    ts = datetime.datetime.now().strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
    print("%s, 0, 1965, 1965, 455.10, Not Active, Not Active, Not Active, Not Active" % ts, flush=True)
    print("%s, 1, 1410, 1965, 700.00, Not Active, Active, Not Active, Active" % ts, flush=True)
    time.sleep(0.02)
'''


def test_clock_sampler_window_and_reasons(tmp_path, monkeypatch):
    monkeypatch.setenv("PATH", _fake_smi(tmp_path, _FAKE) + os.pathsep + os.environ["PATH"])
    s = bench.ClockSampler(0)
    s.start()
    s.wait_ready()
    assert s.lines, "the sampler is live before the warm-up ends"
    time.sleep(0.1)
    before = len(s.lines) // 2
    s.mark_begin()
    time.sleep(0.15)
    s.mark_end()
    time.sleep(0.1)
    out = s.stop()
    assert out["window"] == "timed region"
    assert out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == []
    total = len(s.lines) // 2
    assert 2 <= out["samples"] < total - before + 1, (out, total, before)   # only the window's samples

    # the other GPU's lines are kept apart; throttle reasons come through; a region shorter than
    # the sampling period falls back to the neighbouring samples and says so
    s = bench.ClockSampler(1)
    s.start()
    s.wait_ready()
    time.sleep(0.05)
    s.mark_begin()
    s.mark_end()
    out = s.stop()
    assert out["sm_mhz"] == 1410.0 and out["reasons"] == ["hw_thermal_slowdown", "sw_power_cap"]
    assert out["samples"] >= 1 and out["window"].startswith("timed region +-")


def test_clock_sampler_without_nvidia_smi(tmp_path, monkeypatch):
    monkeypatch.setenv("PATH", str(tmp_path))          # no nvidia-smi anywhere
    s = bench.ClockSampler(0)
    s.start()
    s.wait_ready(timeout=0.1)
    s.mark_begin()
    s.mark_end()
    out = s.stop()
    assert out["sm_mhz"] is None and out["reasons"]


def test_clock_sampler_survives_garbage(tmp_path, monkeypatch):
    body = 'import time\nprint("N/A, 0, [N/A], [N/A], x, a, b, c, d", flush=True)\nprint("junk", flush=True)\ntime.sleep(5)\n'
    monkeypatch.setenv("PATH", _fake_smi(tmp_path, body) + os.pathsep + os.environ["PATH"])
    s = bench.ClockSampler(0)
    s.start()
    s.wait_ready()
    s.mark_begin()
    s.mark_end()
    out = s.stop()
    assert out["sm_mhz"] is None and out["samples"] == 0


def test_committed_traffic_capture_matches_the_kernel_sources():
    """roofline.traffic is quoted only for the sources the ncu capture was taken on
    (profiles/traffic.json carries the hash of csrc/): a kernel edit without a new capture must
    show up here, not as a stale number in the bench line."""
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert t["csrc_sha"] == bench.csrc_sha()
    adj, fwd, src = bench.measured_traffic()
    assert adj == t["adj_kernel_dram_bytes_per_launch"] and fwd == t["fwd_kernel_dram_bytes_per_launch"]
    # the capture is of the kernels the bench names as dominant
    assert "spread_column3d_kernel<float, 6" in t["adj_kernel"]
    assert "interp_fwd_tiled_kernel<float, 3, 6" in t["fwd_kernel"]
    # measured traffic is at least the algorithmic bytes: M(3r + c) + 2 P_K c, P_K c + M(3r + c)
    M, PK = bench.SPOKES * bench.NREAD, 384 ** 3
    assert adj >= M * (3 * 4 + 8) + 2 * PK * 8 and fwd >= PK * 8 + M * (3 * 4 + 8)
