"""CPU checks of the drop-in boundary: the shared library loads, exports every symbol
include/b200nufft.h declares, and the Python binding covers each of them.  No compute."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b200nufft.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2n_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mrrt.nufft_b200 import _lib

    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.load().b2n_version() >= 100


def test_plan_create_argument_errors_without_gpu():
    """Argument validation happens before any CUDA work where possible."""
    from mrrt.nufft_b200 import _lib

    lib = _lib.load()
    plan = ctypes.c_void_p()
    a3 = lambda *v: (ctypes.c_int * 3)(*v)
    rc = lib.b2n_plan_create(4, a3(8, 8, 8), a3(16, 16, 16), a3(6, 6, 6), 1024, 0, 0, 0,
                             ctypes.byref(plan))
    assert rc == _lib.B2N_EINVAL
    assert b"dimensions > 3" in lib.b2n_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)


def test_host_plan_math_matches_golden_tables():
    """Lookup tables are bit-identical to the reference's (float32-accurate in every
    precision), including odd N and complex phasing."""
    from golden_util import tables
    from mrrt.nufft_b200 import _plan_math as pm

    for key, h in tables().items():
        N, K, J, L, ph = key.split("_")
        mine = pm.lookup_table(int(N[1:]), int(J[1:]), int(K[1:]), int(L[1:]), ph)
        assert np.array_equal(mine, h), key


def test_host_plan_math_matches_golden_scaling():
    from golden_util import case_names, load_case
    from mrrt.nufft_b200 import _plan_math as pm
    from mrrt.nufft_b200._kernels import BeattyKernel

    for name in case_names("d[123]_table_*_K32_J6") + case_names("d3_table_double_*") + \
            case_names("d1_table_*") + case_names("d3_mid*"):
        cfg, z = load_case(name)
        Nd, Kd, Jd = cfg["Nd"], cfg["Kd"], cfg["Jd"]
        alphas = [BeattyKernel.beatty_alpha(j, k, n) for j, k, n in zip(Jd, Kd, Nd)]
        sn1d = pm.deapodization_1d(Nd, Kd, Jd, alphas, cfg["phasing"])
        rdt, cdt = pm.real_cplx_dtypes(cfg["precision"])
        assert np.array_equal(pm.dense_sn(sn1d, tuple(Nd)).astype(rdt), z["sn"]), name
        if "phase_after" in z:
            mids = pm.n_mid(Nd, cfg["phasing"])
            pa = pm.phase_after(z["omega"], mids, cfg["n_shift"], rdt, cdt)
            assert np.array_equal(pa, z["phase_after"]), name


def test_no_cpu_fallback():
    import torch
    from mrrt.nufft_b200 import NufftBase

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        NufftBase(Nd=(16, 16), omega=np.zeros((4, 2)), Jd=6)


def test_kernel_objects():
    """tests/test_kernels.py:17-58 of the reference: support, positivity, validation."""
    from mrrt.nufft_b200 import BeattyKernel, kaiser_bessel

    k = BeattyKernel((6, 4), (64, 48), (128, 72))
    for d, J in enumerate((6, 4)):
        x = np.linspace(-J / 2, J / 2, 101)
        y = k.kernels[d](x)
        assert np.all(y[1:-1] > 0) and y[0] == 0 and y[-1] == 0
        assert np.all(k.kernels[d](np.array([J / 2 + 0.1, -J])) == 0)
    assert abs(kaiser_bessel(np.array([0.0]), 6, k.alpha[0])[0] - 1.0) < 1e-15
    with pytest.raises(ValueError):
        BeattyKernel((6, 6), (64,), (128, 128))


def test_kaiser_bessel_matches_reference_values():
    """kaiser_bessel / kaiser_bessel_ft / BeattyKernel against values produced by the real
    reference package (tests/golden/make_golden_kaiser.py) at the reference's own test
    parameters (tests/test_kaiser.py:18-88, tests/test_kernels.py:17-58)."""
    import os
    import warnings

    from mrrt.nufft_b200 import BeattyKernel, kaiser_bessel, kaiser_bessel_ft

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "extra", "kaiser.npz"))
    for m in (-4, 0, 2, 7):
        got = kaiser_bessel(z["kb_x"], 8, 2.34 * 8, m)
        np.testing.assert_allclose(got, z["kb_J8_m%d" % m], rtol=1e-12, atol=1e-14)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for m in (-2, 0, 2, 7):
            np.testing.assert_allclose(kaiser_bessel(z["kbx_x"], 5, 6.8, m), z["kbx_J5_m%d" % m],
                                       rtol=1e-12, atol=1e-14)
            np.testing.assert_allclose(kaiser_bessel_ft(z["ft_u"], 5, 6.8, m, 1), z["ft_J5_m%d" % m],
                                       rtol=1e-11, atol=1e-13)
    shapes = {"c1": ((6, 6), (256, 256), (512, 512)), "c3": ((4, 4, 4), (128,) * 3, (192,) * 3),
              "c5": ((6, 6, 6), (256,) * 3, (384,) * 3), "odd": ((3, 4), (64, 64), (128, 128)),
              "t": ((4, 4), (24, 16), (32, 32))}
    for tag, (shape, grid, os_grid) in shapes.items():
        k = BeattyKernel(shape, grid, os_grid)
        assert np.array_equal(np.asarray(k.alpha, dtype=np.float64), z["beatty_%s_alpha" % tag])
        assert np.array_equal(np.asarray(k.m, dtype=np.float64), z["beatty_%s_m" % tag])
        for d in range(len(shape)):
            xs = np.linspace(-shape[d] / 2 - 0.5, shape[d] / 2 + 0.5, 301)
            np.testing.assert_allclose(k.kernels[d](xs), z["beatty_%s_k%d" % (tag, d)],
                                       rtol=1e-12, atol=1e-14)


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/b200nufft.h compiles as C99 and as C++, and a C program that includes it links
    against libb200nufft.so and gets the documented error code + message for a bad argument
    (validation precedes any CUDA work, so this runs without a GPU)."""
    import shutil
    import subprocess

    from mrrt.nufft_b200 import _lib

    hdr = os.path.join(ROOT, "include", "b200nufft.h")
    gcc, gxx = shutil.which("gcc"), shutil.which("g++")
    if gcc is None or gxx is None:
        pytest.skip("no host compiler")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-x", "c", hdr])
    subprocess.check_call([gxx, "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr])
    src = tmp_path / "use_abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "b200nufft.h"
int main(void) {
    int Nd[3] = {8, 8, 8}, Kd[3] = {16, 16, 16}, Jd[3] = {6, 6, 6};
    b2n_plan *plan = NULL;
    if (b2n_version() < 100) return 1;
    int rc = b2n_plan_create(4, Nd, Kd, Jd, 1024, B2N_SINGLE, 0, 0, &plan);
    if (rc != B2N_EINVAL) return 2;
    if (strstr(b2n_last_error(), "dimensions > 3") == NULL) return 3;
    printf("ok %d\n", b2n_version());
    return 0;
}
''')
    exe = tmp_path / "use_abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-lb200nufft", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("ok"), (out.returncode, out.stdout, out.stderr)
