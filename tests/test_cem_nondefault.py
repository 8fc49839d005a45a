"""Non-default CEM kernels (blurry_cubic_<sigma>, user-supplied ndarray kernels; imresize_CEM.py:22-42):
the oracle and the product's host-side filter derivation against outputs of the unmodified reference
(tests/golden/cem_nondefault.npz, oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from esr_b200 import _capi as capi, cem as pcem
from oracle import cem_filters
from oracle.cem_ops import CEMOracle

CASES = ["blur1_x4", "blur2_x4", "blur07_x2", "blur1_x3", "aniso13_x4", "aniso17s_x4", "aniso15_x2", "aniso15_x3",
         "aniso21s_x4", "aniso13_x4_lmb01"]
OP_CASES = ["blur1_x4", "blur2_x4", "aniso13_x4", "aniso15_x2", "aniso15_x3"]


def kernel_of(g, name):
    k = g[name + "_kernel"]
    return str(k) if k.dtype.kind in "US" else k


@pytest.mark.parametrize("name", CASES)
def test_oracle_filters_match_reference(golden, name):
    g = golden("cem_nondefault")
    f = cem_filters.derive(int(g[name + "_sf"]), kernel=kernel_of(g, name), lower_magnitude_bound=float(g[name + "_lmb"]))
    assert f["ds_kernel"].shape == g[name + "_ds_kernel"].shape and f["inv_hTh"].shape == g[name + "_inv_hTh"].shape
    np.testing.assert_allclose(f["ds_kernel"], g[name + "_ds_kernel"], atol=1e-7)
    np.testing.assert_allclose(f["inv_hTh"], g[name + "_inv_hTh"], atol=2e-6)
    assert [f["margin_LR"], f["margin_HR"], f["ds_half"], f["inv_half"]] == list(g[name + "_margins"])


@pytest.mark.parametrize("name", CASES)
def test_product_filters_match_reference(golden, name):
    g = golden("cem_nondefault")
    sf = int(g[name + "_sf"])
    conf = pcem.Get_CEM_Config(sf)
    conf.lower_magnitude_bound = float(g[name + "_lmb"])      # 0.1 for estimated kernels, SRRaGAN_model.py:63-65
    c = pcem.CEMnet(conf, upscale_kernel=kernel_of(g, name))
    assert c.ds_kernel.shape == g[name + "_ds_kernel"].shape and c.inv_hTh.shape == g[name + "_inv_hTh"].shape
    np.testing.assert_allclose(c.ds_kernel, g[name + "_ds_kernel"], atol=1e-7)
    np.testing.assert_allclose(c.inv_hTh, g[name + "_inv_hTh"], atol=2e-6)
    assert [int(c.invalidity_margins_LR), int(c.invalidity_margins_HR), int(c.ds_kernel_invalidity_half_size_LR),
            int(c.inv_hTh_invalidity_half_size)] == list(g[name + "_margins"])
    # a mild blur keeps both filters rank-1 (separable kernels); a strong blur or an estimated kernel does not
    assert c.separable == (name in ("blur1_x4", "blur07_x2", "blur1_x3"))
    if c.separable:
        np.testing.assert_allclose(np.outer(c._ds_1d, c._ds_1d), g[name + "_ds_kernel"], atol=1e-7)
        np.testing.assert_allclose(np.outer(c._inv_1d, c._inv_1d), g[name + "_inv_hTh"], atol=2e-6)
        assert isinstance(c._filters, capi.CemFilters)
    else:
        assert isinstance(c._filters, capi.CemFilterBank2D)
        assert c._filters.ds.shape == c.ds_kernel.shape and c._filters.inv.dtype == np.float32


@pytest.mark.parametrize("name", OP_CASES)
def test_oracle_ops_match_reference(golden, name):
    g = golden("cem_nondefault")
    sf = int(g[name + "_sf"])
    ora = CEMOracle(sf, filters=cem_filters.derive(sf, kernel=kernel_of(g, name)))
    y, x = torch.from_numpy(g[name + "_y"]), torch.from_numpy(g[name + "_x"])
    np.testing.assert_allclose(ora.downscale(y).numpy(), g[name + "_down"], atol=2e-6)
    np.testing.assert_allclose(ora.upscale(x).numpy(), g[name + "_up"], atol=2e-6)
    np.testing.assert_allclose(ora.conv_inv_hTh(x).numpy(), g[name + "_inv"], atol=1e-5)
    np.testing.assert_allclose(ora.project(y, x).numpy(), g[name + "_project"], atol=2e-5)
    yg = y.clone().requires_grad_(True)
    (ora.project(yg, x) * torch.from_numpy(g[name + "_grad_g"])).sum().backward()
    np.testing.assert_allclose(yg.grad.numpy(), g[name + "_grad_y"], atol=2e-5)


def test_kernel_argument_errors():
    """Same refusals as the reference: unknown names, kernels that do not sum to 1 (imresize_CEM.py:8,26)."""
    with pytest.raises(ValueError):
        pcem.CEMnet(pcem.Get_CEM_Config(4), upscale_kernel="lanczos")
    with pytest.raises(ValueError):
        pcem.CEMnet(pcem.Get_CEM_Config(4), upscale_kernel=np.ones((9, 9)))
    with pytest.raises(AssertionError):
        pcem.CEMnet(pcem.Get_CEM_Config(4), upscale_kernel=7)
    # filters wider than the stencil kernels accept are refused on the host, before any launch
    with pytest.raises(capi.EsrError):
        capi.CemFilterBank2D(4, 1, np.zeros((65, 65), np.float32), np.zeros((5, 5), np.float32))
    with pytest.raises(capi.EsrError):
        capi.CemFilterBank2D(4, 1, np.zeros((9, 9), np.float32), np.zeros((6, 6), np.float32))


def test_c_abi_rejects_bad_2d_filters_before_any_launch():
    """esr_cem2d_* validate the filter descriptor first (no CUDA call has happened yet): even / oversized sides, null
    tap pointers and unsupported scale factors come back as an error code with a message, on a box without a GPU too."""
    import ctypes as C
    l = capi.lib()
    buf = (C.c_float * 16)()
    addr = C.cast(buf, C.c_void_p).value
    for sf, n_ds, n_inv, ds, inv, needle in ((4, 8, 5, addr, addr, "ds kernel side"), (4, 9, 65, addr, addr, "inv_hTh side"),
                                             (4, 9, 5, 0, addr, "null"), (5, 9, 5, addr, addr, "scale factor")):
        f = capi.CemFilters2d()
        f.sf, f.pre, f.n_ds, f.n_inv, f.ds, f.inv = sf, 1, n_ds, n_inv, ds, inv
        rc = l.esr_cem2d_downscale(f, addr, 1, 1, 8, 8, addr, None)
        assert rc < 0
        assert needle in l.esr_last_error().decode()
        with pytest.raises(capi.EsrError):
            capi.check(l.esr_cem2d_project_bwd(f, addr, 1, 1, 8, 8, 0, addr, addr, None))
