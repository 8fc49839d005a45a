"""Z optimisation (BASELINE config 3 path) against loss trajectories recorded from the reference's own
``Z_optimizer`` (tests/golden/zopt.npz, oracle/gen_golden.py) - same objective, Adam steps and inputs."""
import numpy as np
import pytest
import torch

from esr_b200 import synth
from esr_b200.z_optimization import Z_optimizer, SRModelShim
from tests.test_gpu_net import build_product_G

pytestmark = pytest.mark.gpu


def _setup(dev, train_mode):
    wts = synth.make_weights("default", seed=5, nb=2)
    netG = build_product_G(dev, 2, "all_layers_HR_downscaled", wts, train=train_mode)
    lr, z0 = synth.make_inputs(1, 8, 8, seed=5)
    model = SRModelShim(netG)
    data = {"LR": lr.to(dev), "Z": (0.5 * z0).to(dev)}
    model.feed_data(data)
    with torch.no_grad():
        model.fake_H = netG(model.model_input)
    return netG, model, data


def test_tv_objective_gui_mode(golden, cuda_device):
    g = golden("zopt")
    netG, model, data = _setup(cuda_device, train_mode=False)
    opt = Z_optimizer(objective="TV", Z_size=[32, 32], model=model, Z_range=1.0, max_iters=4, data=data, initial_LR=0.1,
                      batch_size=1)
    Z = opt.optimize()
    np.testing.assert_allclose(np.array(opt.loss_values), g["tv_eval_loss"], rtol=3e-3)
    assert float((Z.cpu() - torch.from_numpy(g["tv_eval_Z"])).abs().mean()) < 2e-2
    assert all(not p.requires_grad for p in netG.parameters())
    assert opt.loss_values[-1] < opt.loss_values[0]


def test_l1_objective_training_mode(golden, cuda_device):
    g = golden("zopt")
    netG, model, data = _setup(cuda_device, train_mode=True)
    del model.__dict__["fake_H"]
    data["HR"] = torch.from_numpy(g["l1_train_target"]).to(cuda_device)
    opt = Z_optimizer(objective="l1", Z_size=[32, 32], model=model, Z_range=1.0, max_iters=4, data=data, initial_LR=0.1,
                      batch_size=1, HR_unpadder=lambda t: t)
    opt.feed_data(data)
    opt.random_Z_inits = False
    opt.Z_model.Z.data.copy_(torch.from_numpy(g["l1_train_Zinit"]).to(cuda_device))
    Z = opt.optimize()
    np.testing.assert_allclose(np.array(opt.loss_values), g["l1_train_loss"], rtol=3e-3)
    assert float((Z.cpu() - torch.from_numpy(g["l1_train_Z"])).abs().mean()) < 2e-2


def test_unbuilt_objective_is_loud(cuda_device):
    netG, model, data = _setup(cuda_device, train_mode=False)
    with pytest.raises(NotImplementedError):
        Z_optimizer(objective="hist", Z_size=[32, 32], model=model, Z_range=1.0, max_iters=4, data=data, initial_LR=0.1)
