"""Pins the CPU oracle against outputs of the unmodified reference (tests/golden/*.npz, produced by
oracle/gen_golden.py in the build container).  CPU only."""
import numpy as np
import pytest
import torch

from esr_b200 import synth
from oracle import cem_filters
from oracle.cem_ops import CEMOracle, concat_latent
from oracle.rrdbnet import GCEMOracle


@pytest.mark.parametrize("sf", [2, 3, 4])
def test_filters_match_reference(golden, sf):
    g = golden("cem_filters")
    f = cem_filters.derive(sf)
    assert f["ds_kernel"].shape == g["ds_kernel_%d" % sf].shape
    np.testing.assert_allclose(f["ds_kernel"], g["ds_kernel_%d" % sf], atol=1e-7)
    np.testing.assert_allclose(f["inv_hTh"], g["inv_hTh_%d" % sf], atol=2e-6)
    assert [f["margin_LR"], f["margin_HR"], f["ds_half"], f["inv_half"]] == list(g["margins_%d" % sf])


def test_bicubic_x4_known_taps():
    """SURVEY.md §8(a16) known-answer values."""
    f = cem_filters.derive(4)
    r = f["ds_kernel"][8]  # centre row = d[8] * d
    d = np.sqrt(np.diag(f["ds_kernel"]).clip(0))
    np.testing.assert_allclose(d[7], 0.2418212891, atol=1e-7)
    assert f["ds_kernel"].shape == (17, 17) and np.all(f["ds_kernel"][-1] == 0) and np.all(f["ds_kernel"][:, -1] == 0)
    assert np.linalg.matrix_rank(f["ds_kernel"].astype(np.float64), tol=1e-9) == 1
    assert f["inv_hTh"].shape == (27, 27)
    np.testing.assert_allclose(np.sqrt(f["inv_hTh"][13, 13]), 1.2469131186, atol=2e-6)
    assert (f["margin_LR"], f["margin_HR"]) == (10, 40)
    assert r.shape == (17,)


def test_cem_ops_match_reference(golden):
    g = golden("cem_ops")
    cem = CEMOracle(4)
    y, x = torch.from_numpy(g["y"]), torch.from_numpy(g["x"])
    np.testing.assert_allclose(cem.downscale(y).numpy(), g["down"], atol=2e-6)
    np.testing.assert_allclose(cem.upscale(x).numpy(), g["up"], atol=2e-6)
    np.testing.assert_allclose(cem.conv_inv_hTh(x).numpy(), g["inv"], atol=2e-6)
    np.testing.assert_allclose(cem.project(y, x).numpy(), g["project_train"], atol=1e-5)
    yg = y.clone().requires_grad_(True)
    (cem.project(yg, x) * torch.from_numpy(g["project_grad_g"])).sum().backward()
    np.testing.assert_allclose(yg.grad.numpy(), g["project_grad_y"], atol=1e-5)
    mi = torch.from_numpy(g["eval_model_input"])
    xp = cem.pre_pad(mi, 3)
    out = cem.unpad_HR(cem.project(torch.from_numpy(g["eval_y"]), xp[:, -3:]))
    np.testing.assert_allclose(out.numpy(), g["project_eval"], atol=1e-5)


CASES = ["prod_default", "prod_kaiming", "nb2_train_mode", "nb2_first_layer", "nb1_no_latent"]


def oracle_for_case(g, name, operand_dtype=None):
    nb, seed, B, h, w, train = [int(v) for v in g[name + "_cfg"]]
    li, kind = [str(v) for v in g[name + "_str"]]
    latent = None if li == "None" else li + "_HR_downscaled"
    wts = synth.make_weights(kind, seed=seed, nb=nb, latent_input=latent)
    lr, z = synth.make_inputs(B, h, w, seed=seed)
    mi = lr if latent is None else concat_latent(lr, z)
    net = GCEMOracle(wts, pre_pad=not train, nb=nb, latent_input=latent,
                     num_latent_channels=3 if latent else 0, operand_dtype=operand_dtype)
    return net, mi, wts, dict(nb=nb, latent=latent, train=bool(train), kind=kind, seed=seed)


@pytest.mark.parametrize("name", CASES)
def test_g_cem_matches_reference(golden, name):
    g = golden("g_cem")
    net, mi, _, _ = oracle_for_case(g, name)
    with torch.no_grad():
        out = net.forward(mi)
    np.testing.assert_allclose(out.numpy(), g[name + "_out"], atol=2e-5)


def test_param_count_and_flops():
    """17 060 948 trainable parameters; §8: 18 316 944 MAC per LR pixel."""
    shapes = synth.rrdbnet_conv_shapes()
    assert sum(co * ci * 9 + co for co, ci in shapes.values()) == 17060948
    assert len(shapes) == 351
    macs = 0
    for k, (co, ci) in shapes.items():
        res = 1 if k.startswith(("model.0", "model.1")) else (4 if k.startswith("model.2") else 16)
        macs += 9 * co * ci * res
    assert macs == 18316944


X2_CASES = ["x2_nb2_eval", "x2_nb1_train"]


@pytest.mark.parametrize("name", X2_CASES)
def test_x2_generator_matches_reference(golden, name):
    """x2 (one nearest-x2 upconv stage): forward and the Z gradient of the reference's autograd."""
    g = golden("g_cem_x2")
    nb, seed, B, h, w, train = [int(v) for v in g[name + "_cfg"]]
    wts = synth.make_weights(str(g[name + "_kind"]), seed=seed, nb=nb, upscale=2)
    lr, z = synth.make_inputs(B, h, w, sf=2, seed=seed)
    ora = GCEMOracle(wts, sf=2, pre_pad=not train, nb=nb)
    zg = z.clone().requires_grad_(True)
    out = ora.forward(concat_latent(lr, zg, sf=2))
    np.testing.assert_allclose(out.detach().numpy(), g[name + "_out"], atol=2e-5)
    (out * torch.from_numpy(g[name + "_gout"])).sum().backward()
    np.testing.assert_allclose(zg.grad.numpy(), g[name + "_gz"], atol=2e-5)


@pytest.mark.parametrize("name", ["lr_all_nb2_train", "lr_all_nb1_eval", "lr_first_nb1_train"])
def test_lr_domain_plumbing_maps_onto_the_hr_downscaled_path(golden, name):
    """The product serves ``latent_input_domain: "LR"`` by feeding its HR_downscaled path with
    Z_HR = nearest_upsample(Z, 4) (rrdbnet._lr_domain_input).  Here that host-side packing (no kernels involved) goes
    through the CPU oracle of the HR_downscaled path and must reproduce the reference's LR-domain output and Z gradient."""
    from esr_b200.rrdbnet import RRDBNet, _lr_domain_input
    g = golden("g_cem_lr_domain")
    nb, seed, train = [int(v) for v in g[name + "_cfg"]]
    latent = str(g[name + "_latent"])
    wts = synth.make_weights(str(g[name + "_kind"]), seed=seed, nb=nb, latent_input=latent + "_HR_downscaled")
    net = RRDBNet(3, 3, 64, nb, latent_input=latent + "_LR", num_latent_channels=3)
    assert [k for k, _ in net.named_parameters()] == list(wts)
    z = torch.from_numpy(g[name + "_z"]).requires_grad_(True)
    net.Z = z
    margin = 0 if train else 10
    packed = _lr_domain_input(net, torch.from_numpy(g[name + "_lr"]), margin)
    assert packed.shape[1] == 51 and packed.shape[2:] == z.shape[2:]
    ora = GCEMOracle(wts, pre_pad=False, nb=nb, latent_input=latent + "_HR_downscaled", num_latent_channels=3)
    out = ora.forward(packed)
    c = 4 * margin
    out = out[..., c:out.size(-2) - c, c:out.size(-1) - c]
    np.testing.assert_allclose(out.detach().numpy(), g[name + "_out"], atol=2e-5)
    (out * torch.from_numpy(g[name + "_gout"])).sum().backward()
    np.testing.assert_allclose(z.grad.numpy(), g[name + "_gz"], atol=1e-5 * float(np.abs(g[name + "_gz"]).max()) + 1e-7)


def test_hr_rearranged_domain_is_refused():
    from esr_b200.rrdbnet import RRDBNet
    with pytest.raises(NotImplementedError, match="HR_rearranged"):
        RRDBNet(3, 3, 64, 1, latent_input="all_layers_HR_rearranged", num_latent_channels=3)
