"""Rich Z objectives on the B200: the density kernels of csrc/zobj.cu against the oracle, the
product's SoftHistogramLoss against values / gradients recorded from the unmodified reference, and the objectives through
Z_optimizer around the real G+CEM against the reference's Z_optimizer trajectories (tests/golden/zobjectives.npz,
oracle/gen_golden.py: gen_zobj)."""
import warnings

import numpy as np
import pytest
import torch

from esr_b200 import synth, z_objectives as zo
from esr_b200.z_optimization import Z_optimizer, SRModelShim
from oracle import zobj_cases as zc, zobjectives as oracle_zobj
from tests.test_gpu_net import build_product_G

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore")


@pytest.mark.parametrize("D,N,M,temperature", [(1, 1000, 256, 5e-4),        # gray histogram: few bins, the sample range is split
                                               (1, 70001, 256, 5e-4),
                                               (36, 333, 517, 1e-3),        # 6x6 patches against a patch dictionary
                                               (49, 129, 31, 2e-3),
                                               (3, 4099, 65, 1e-2),
                                               (81, 40, 40, 5e-3)])         # the largest supported sample (9x9)
def test_density_sums_and_gradients_match_oracle(cuda_device, D, N, M, temperature):
    """esr_kde_sums / esr_kde_grad against the fp64 restatement of Z_optimization.py:184-195, per sample and per bin, with
    the wrap-around distance active (samples and bins spread over the whole period)."""
    rng = np.random.default_rng(D * 1000 + N)
    base = rng.random((D, 1))
    spread = 0.08 if D > 1 else 1.0                      # high-dimensional samples must stay close or everything underflows
    samples = torch.from_numpy(((base + spread * rng.random((D, N))) % 1.0).astype(np.float32))
    bins = torch.from_numpy((base + spread * rng.random((D, M))) % 1.0)
    for per_bin in (False, True):
        x_ref = samples.clone().requires_grad_(True)
        ref = oracle_zobj.kde_sums(x_ref, bins, 1.0, temperature, per_bin=per_bin)
        weight = torch.from_numpy(rng.random(ref.shape[0]))
        (ref * weight).sum().backward()
        x = samples.to(cuda_device).requires_grad_(True)
        out = zo.kde_sums(x, bins.to(cuda_device), 1.0, temperature, per_bin=per_bin)
        assert out.dtype == torch.float64 and float(ref.max()) > 1e-6        # the case is not an all-underflow one
        np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-9, atol=1e-300)
        (out * weight.to(cuda_device)).sum().backward()
        g_ref = x_ref.grad.numpy()
        np.testing.assert_allclose(x.grad.cpu().numpy(), g_ref, rtol=2e-5, atol=1e-6 * np.abs(g_ref).max())
        again = zo.kde_sums(samples.to(cuda_device), bins.to(cuda_device), 1.0, temperature, per_bin=per_bin)
        assert torch.equal(again, out.detach())                              # fixed summation order


def test_density_kernel_argument_errors(cuda_device):
    from esr_b200 import _capi as capi
    x, b = torch.rand(2, 16, device=cuda_device), torch.rand(2, 8, device=cuda_device).double()
    with pytest.raises(capi.EsrError):
        zo.kde_sums(x, b, 1.0, 0.0)                                          # temperature must be positive
    with pytest.raises(capi.EsrError):
        zo.kde_sums(torch.rand(82, 16, device=cuda_device), torch.rand(82, 8, device=cuda_device).double(), 1.0, 1e-3)


@pytest.mark.parametrize("name", sorted(zc.HIST_CASES))
def test_soft_histogram_loss_matches_reference_golden(golden, cuda_device, name):
    g = golden("zobjectives")
    value, grad, bins = zc.run_hist_case(zo.SoftHistogramLoss, name, device=cuda_device)
    ref_bins = g["hist_%s_bins" % name]                # same atoms kept by the pruning; the GPU's channel mean differs by 1 ulp
    assert bins.shape == ref_bins.shape
    np.testing.assert_allclose(bins.numpy(), ref_bins, rtol=0, atol=1e-6)
    np.testing.assert_allclose(value.numpy().astype(np.float64), g["hist_%s_value" % name], rtol=2e-5, atol=1e-7)
    ref = g["hist_%s_grad" % name]
    np.testing.assert_allclose(grad.numpy(), ref, rtol=1e-3, atol=2e-5 * np.abs(ref).max())


# per case: relative tolerance of the loss trajectory.  The generator runs bf16 tensor-core arithmetic, so the output
# differs from the fp32 reference by ~1e-3; squared STD differences and KL divergences of nearly equal histograms are
# differences of nearly equal numbers and amplify that.
ZOPT3_RTOL = {"local_std_up": 5e-2, "local_std_max": 5e-3, "local_mag_up": 5e-2, "local_tv": 5e-3, "periodicity_int": 5e-3,
              "periodicity_nonint": 5e-3, "periodicity_plus": 2e-2, "random_l1": 2e-2, "random_l1_limited": 5e-2,
              "hist_keepstd": 0.25, "patchhist_noDC_keepstd": 0.1, "dict": 1e-3, "patchdict_noDC": 2e-2,
              "patchdict_noDC_keepstd": 2e-2}


@pytest.mark.parametrize("name", [n for n in zc.ZOPT3_CASES if n not in zc.NO_REFERENCE_RUN])
def test_objectives_through_the_generator_match_reference_trajectories(golden, cuda_device, name):
    g = golden("zobjectives")
    dev = cuda_device
    netG = build_product_G(dev, 2, "all_layers_HR_downscaled", synth.make_weights("default", seed=5, nb=2), train=False)
    lr, z0 = synth.make_inputs(1, zc.ZOPT3_HW[0], zc.ZOPT3_HW[1], seed=11)
    opt, Z = zc.run_zopt_case(Z_optimizer, SRModelShim(netG), netG, name, lr, z0, device=dev, z_init=zc.zopt3_z_init(name))
    ref = g["zopt_%s_loss" % name]
    assert len(opt.loss_values) == len(ref)
    rtol = ZOPT3_RTOL[name]
    dev_loss = float(np.max(np.abs(np.array(opt.loss_values) - ref) / np.maximum(np.abs(ref), 1e-30)))
    dev_z = float((Z.cpu() - torch.from_numpy(g["zopt_%s_Z" % name])).abs().mean())
    dev_std = float(np.max(np.abs(opt.initial_STD.cpu().numpy() - g["zopt_%s_initial_STD" % name])))
    print("ZOPT3 %s: worst relative loss deviation %.3g (bound %g), mean |dZ| %.3g, worst initial-STD deviation %.3g" % (name, dev_loss, rtol, dev_z, dev_std))
    np.testing.assert_allclose(opt.initial_STD.cpu().numpy(), g["zopt_%s_initial_STD" % name], rtol=5e-3, atol=2e-3)
    np.testing.assert_allclose(np.array(opt.loss_values), ref, rtol=rtol, atol=rtol * np.abs(ref).max())
    np.testing.assert_allclose(np.array(opt.latest_Z_loss_values).reshape(-1), g["zopt_%s_latest" % name], rtol=rtol,
                               atol=rtol * np.abs(ref).max())
    # Adam's steps are lr-sized whatever the gradient's magnitude: where the objective's gradient is weak (the plain
    # dictionary over 256 uniform bins: ~1e-8) the bf16 noise of G's backward decides the step's sign
    # (measured on a B200, gpurun_out/zopt3_margins.log: <= 7e-3 everywhere except the two histogram cases, whose STD-preserving
    # term with weight 1e4 makes the first steps overshoot - the loss GROWS 2e-4 -> 0.08 in the reference too: 0.025 / 0.010)
    z_bound = {"dict": 0.15, "hist_keepstd": 0.1, "patchhist_noDC_keepstd": 0.06}.get(name, 3e-2)
    assert dev_z < z_bound


def test_scribble_objective_runs_on_the_generator(cuda_device):
    """No reference run exists for 'scribble' (zobj_cases.NO_REFERENCE_RUN; its arithmetic is pinned on CPU by
    tests/test_zobjectives.py): here it only has to drive the real generator's Z downhill."""
    dev = cuda_device
    netG = build_product_G(dev, 2, "all_layers_HR_downscaled", synth.make_weights("default", seed=5, nb=2), train=False)
    lr, z0 = synth.make_inputs(1, zc.ZOPT3_HW[0], zc.ZOPT3_HW[1], seed=11)
    opt, Z = zc.run_zopt_case(Z_optimizer, SRModelShim(netG), netG, "scribble", lr, z0, device=dev)
    assert len(opt.loss_values) == 3 and opt.loss_values[-1] < opt.loss_values[0]
    assert torch.isfinite(Z).all()
