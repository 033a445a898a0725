"""Executes the reference-side binding of INTEGRATION.md section 2 (integration/reference_stub.py:
``init_gpu`` / ``table_interp`` / ``table_adj`` through the RAW C ABI with ``B2N_COORD_TM``
coordinates and the reference's own ``h`` tables) on reference-side operator objects, and a
DLPack producer round trip through ``NufftBase``.  The "reference operator" is the oracle's
restatement of ``NufftBase`` (same attributes: ndim, Nd, Kd, Jd, Ld, precision, phasing, h, tm,
M); CuPy is absent from this image, so the device-array type is PyTorch (``TorchArrays``)."""
import os
import sys

import numpy as np
import pytest

from golden_util import TOL, grid_only_inputs, load_case, rel_l2, ctor_kwargs

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "integration"))


@pytest.mark.parametrize("name", ["d1_table_single_real", "d1_table_double_complex",
                                  "d2_table_single_real_K33_J7", "d2_radial_table_single_real",
                                  "d3_table_double_real", "d3_mid_table_single_real_J546",
                                  "d3_table_double_complex"])
def test_reference_side_stub(name):
    import reference_stub as stub
    from oracle import nufft_oracle as orc

    cfg, z = load_case(name)
    kw = ctor_kwargs(cfg)
    O = orc.OracleNufft(omega=z["omega"], **kw)        # the reference-side operator object
    assert np.array_equal(O.tm, z["tm"])
    xp = stub.TorchArrays()
    stub.init_gpu(O, xp)
    try:
        g, ysamp = grid_only_inputs(cfg["seed"], int(np.prod(O.Kd)), O.M, cfg["n_reps"], O._cplx_dtype)
        tol = TOL[cfg["precision"]]
        out = stub.table_interp(O, g, xp)
        assert tuple(out.shape) == (O.M, cfg["n_reps"])
        want = orc.interp_table(O.Kd, O.Jd, O.Ld, O.h, O.tm, g)      # no phase_shift: the stub
        assert rel_l2(out.cpu().numpy(), want) <= tol                # replaces the kernel launch only
        gk = stub.table_adj(O, ysamp, xp)
        assert tuple(gk.shape) == (int(np.prod(O.Kd)), cfg["n_reps"])
        want = orc.interp_table_adj(O.Kd, O.Jd, O.Ld, O.h, O.tm, ysamp)
        assert rel_l2(gk.cpu().numpy(), want) <= tol
        # ... and it reproduces what the reference package itself produced where its
        # interpolation stage has no phase_shift (real phasing, or n_shift == 0)
        if cfg["phasing"] == "real" or not any(cfg["n_shift"]):
            assert rel_l2(out.cpu().numpy(), z["interp_out"]) <= tol
    finally:
        stub.destroy(O)


class _Producer(object):
    """A bare DLPack producer (neither torch nor CuPy): what ``_ArrayKind`` calls "dlpack"."""

    def __init__(self, t):
        self._t = t

    def __dlpack__(self, stream=None, **kw):
        return self._t.__dlpack__(stream=stream) if stream is not None else self._t.__dlpack__()

    def __dlpack_device__(self):
        return self._t.__dlpack_device__()


def test_dlpack_producer_round_trip():
    import torch
    from mrrt.nufft_b200 import NufftBase

    cfg, z = load_case("d2_radial_table_single_real")
    A = NufftBase(omega=z["omega"], on_gpu=True, **ctor_kwargs(cfg))
    xt = torch.from_numpy(np.asfortranarray(z["x"])).cuda()
    y = A.fft(_Producer(xt))
    # same device, zero host copies: a DLPack producer comes back
    assert hasattr(y, "__dlpack__") and y.is_cuda
    assert rel_l2(torch.from_dlpack(y).cpu().numpy(), z["y"]) <= TOL["single"]
    xa = A.adj(_Producer(torch.from_numpy(np.asfortranarray(z["y"])).cuda()))
    assert hasattr(xa, "__dlpack__") and xa.is_cuda
    assert rel_l2(torch.from_dlpack(xa).cpu().numpy(), z["x_adj"]) <= TOL["single"]
    # the consumer did not copy: the tensor made from the capsule aliases the producer's memory
    assert torch.from_dlpack(_Producer(xt)).data_ptr() == xt.data_ptr()
