"""Generator weight gradients (csrc/wgrad.cu, training.GeneratorTrainer) against autograd through the CPU oracle: the
generator half of the reference's training step (codes/models/SRRaGAN_model.py:463-547 with trainable G parameters).
Tolerance: bf16 gradient and activation operands (8 significant bits, random rounding errors averaged over the pixels):
relative error of every weight gradient < 8 % (bias gradient < 10 %), cosine > 0.996 (0.995); measured 2.6-6.5 % (7.8 %).
Most of it is LeakyReLU masks decided on the 16-bit forward activations: about 0.4 % of them sit within rounding of zero
and take slope 1 instead of 0.2 or vice versa, which alone is sqrt(0.004) x 0.8 = 5 % of the gradient norm.  The error is the data-gradient chain's (the bound of
tests/test_gpu_backward.py): measured 2.6-4.8 %, equal for the latent (exact fp32 input) and the 16-bit input channels of
a conv, largest for the earliest blocks whose gradient went through the most bf16 dgrads."""
import pytest
import torch

from esr_b200 import synth
from esr_b200.training import GeneratorTrainer
from oracle.cem_ops import concat_latent
from oracle.rrdbnet import GCEMOracle
from tests.test_gpu_net import build_product_G

pytestmark = pytest.mark.gpu


def _oracle_grads(wts, nb, mi, gout, train):
    w = {k: v.clone().requires_grad_(True) for k, v in wts.items()}
    out = GCEMOracle(w, pre_pad=not train, nb=nb).forward(mi)
    (out * gout).sum().backward()
    return out.detach(), {k: v.grad for k, v in w.items()}


@pytest.mark.parametrize("nb,B,h,w,train,kind", [(1, 2, 24, 32, True, "default"), (2, 1, 28, 24, False, "kaiming"),
                                                 (23, 1, 32, 32, True, "kaiming")])     # BASELINE config 5's generator and patch size
def test_weight_gradients_match_oracle_autograd(cuda_device, nb, B, h, w, train, kind):
    wts = synth.make_weights(kind, seed=3, nb=nb)
    lr, z = synth.make_inputs(B, h, w, seed=3)
    mi = concat_latent(lr, z)
    netG = build_product_G(cuda_device, nb, "all_layers_HR_downscaled", wts, train=train)
    trainer = GeneratorTrainer(netG)
    fake = trainer.forward(mi.to(cuda_device))
    gout = torch.randn(fake.shape, generator=torch.Generator().manual_seed(4))
    ref_out, ref = _oracle_grads(wts, nb, mi, gout, train)
    assert (fake.detach().cpu() - ref_out).abs().max().item() <= 1e-2
    trainer.backward(gout.to(cuda_device))
    G = netG.generated_image_model
    worst = (0.0, None)
    for name, p in G.named_parameters():
        got, want = p.grad.cpu(), ref[name]
        assert got.shape == want.shape and torch.isfinite(got).all(), name
        rel = float((got - want).norm() / want.norm().clamp_min(1e-20))
        cos = float((got * want).sum() / (got.norm() * want.norm()).clamp_min(1e-30))
        worst = max(worst, (rel, name))
        if want.dim() == 4 and want.shape[1] > 3:
            print("%-40s rel %.4f cos %.5f | latent part rel %.4f, main part rel %.4f" % (
                name, rel, cos, float((got[:, :3] - want[:, :3]).norm() / want[:, :3].norm()), float((got[:, 3:] - want[:, 3:]).norm() / want[:, 3:].norm())))
        if name == "model.6.bias":
            # the CEM projection does not depend on a constant added to the generator's output, so this gradient is zero up
            # to border effects: compare absolutely (a sum of bf16-rounded terms would miss by ~0.5 here, fp32 sums by 1e-4)
            assert float((got - want).abs().max()) <= 1e-6 * float(gout.abs().sum()), name
            continue
        # a bias gradient is the plain sum of the bf16 gradient over the pixels (signs cancel): a little above the weights'
        lim, cmin = (10e-2, 0.995) if name.endswith(".bias") else (8e-2, 0.996)
        assert rel < lim and cos > cmin, "%s: relative error %g, cosine %g" % (name, rel, cos)
    print("worst relative error", worst)


def test_training_steps_reduce_the_loss_and_update_in_place(cuda_device):
    """Three Adam steps of an L1 pixel loss through GeneratorTrainer: the loss falls, and the weight updates reach the
    kernels without rebuilding plans (the packed weight images are refreshed inside their buffers)."""
    wts = synth.make_weights("kaiming", seed=6, nb=1)
    lr, z = synth.make_inputs(2, 16, 16, seed=6)
    netG = build_product_G(cuda_device, 1, "all_layers_HR_downscaled", wts, train=True)
    G = netG.generated_image_model
    for p in G.parameters():
        p.requires_grad_(True)
    trainer = GeneratorTrainer(netG)
    opt = torch.optim.Adam(G.parameters(), lr=1e-3)
    target = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(7)).to(cuda_device)
    mi = concat_latent(lr, z).to(cuda_device)
    losses, plans = [], []
    for _ in range(3):
        fake = trainer.forward(mi)
        loss = (fake - target).abs().mean()
        loss.backward()
        trainer.backward(fake.grad)
        opt.step()
        losses.append(float(loss))
        plans.append(id(trainer._bp[1]))
    assert losses[2] < losses[0], losses
    assert len(set(plans)) == 1, "the backward plan was rebuilt after a weight update"


def test_autograd_node_returns_weight_gradients(cuda_device):
    """``netG(model_input)`` with trainable parameters (the reference's own training call, SRRaGAN_model.py:349,533):
    ``loss.backward()`` fills ``p.grad`` through autograd with the same values as the explicit GeneratorTrainer, a
    second forward + backward accumulates, and the latent gradient comes back too."""
    nb, B, h, w = 1, 2, 16, 24
    wts = synth.make_weights("kaiming", seed=8, nb=nb)
    lr, z = synth.make_inputs(B, h, w, seed=8)
    mi = concat_latent(lr, z).to(cuda_device)
    ref_net = build_product_G(cuda_device, nb, "all_layers_HR_downscaled", wts, train=True)
    trainer = GeneratorTrainer(ref_net)
    fake = trainer.forward(mi)
    gout = torch.randn(fake.shape, generator=torch.Generator().manual_seed(9)).to(cuda_device)
    g_in_ref = trainer.backward(gout).clone()
    want = {k: p.grad.clone() for k, p in ref_net.generated_image_model.named_parameters()}

    netG = build_product_G(cuda_device, nb, "all_layers_HR_downscaled", wts, train=True)
    for p in netG.parameters():
        p.requires_grad_(True)
    x = mi.clone().requires_grad_(True)
    out = netG(x)
    assert out.requires_grad and torch.equal(out.detach(), fake.detach())
    (out * gout).sum().backward()
    assert torch.allclose(x.grad, g_in_ref, rtol=1e-4, atol=1e-6 * float(g_in_ref.abs().max()))
    for k, p in netG.generated_image_model.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape, k
        # chunked high-resolution items add with atomics: equal up to fp32 summation order
        assert float((p.grad - want[k]).norm()) <= 1e-4 * float(want[k].norm()) + 1e-7 * float(gout.abs().sum()), k
    (netG(mi) * gout).sum().backward()                       # accumulation, like any torch module
    for k, p in netG.generated_image_model.named_parameters():
        assert float((p.grad - 2 * want[k]).norm()) <= 2e-4 * float(want[k].norm()) + 2e-7 * float(gout.abs().sum()), k
    out2 = netG(mi)
    netG(mi)
    with pytest.raises(Exception, match="one forward per backward"):
        (out2 * gout).sum().backward()
    # frozen generator, grad enabled (the reference's D-only steps): plain forward, nothing recorded
    for p in netG.parameters():
        p.requires_grad_(False)
    assert not netG(mi).requires_grad


@pytest.mark.parametrize("crop", [0, 8])
def test_gan_step_matches_torch_autograd_on_the_oracle(cuda_device, crop):
    """training.GanTrainer.step (BASELINE config 5; SRRaGAN_model.py:349-547 in the shipped wgan-gp configuration)
    against the same iteration written with torch autograd on the CPU oracle generator and a CPU copy of the critic:
    critic losses incl. the gradient penalty, the critic's Adam update, the generator's loss terms, its weight gradients
    (tolerance as above plus the critic's BatchNorm in the chain) and its Adam update."""
    import copy
    from esr_b200.discriminator import Discriminator_VGG_128_
    from esr_b200.training import GanTrainer
    nb, B, h = 1, 2, 24
    wts = synth.make_weights("kaiming", seed=11, nb=nb)
    lr, z = synth.make_inputs(B, h, h, seed=11)
    mi = concat_latent(lr, z)
    hr = torch.rand(B, 3, 4 * h, 4 * h, generator=torch.Generator().manual_seed(12))
    u = torch.rand(B, 1, 1, 1, generator=torch.Generator().manual_seed(13))
    torch.manual_seed(14)
    d_cpu = Discriminator_VGG_128_(3, 8, nb=4, input_patch_size=4 * h - 2 * crop)
    for m in d_cpu.modules():
        if isinstance(m, torch.nn.Conv2d):
            torch.nn.init.kaiming_normal_(m.weight, a=0, mode='fan_in')
    d_gpu = copy.deepcopy(d_cpu).to(cuda_device)
    weights = dict(pixel_weight=1e-2, gan_weight=5e-3, gp_weight=10.0, range_weight=50.0)
    netG = build_product_G(cuda_device, nb, "all_layers_HR_downscaled", wts, train=True)
    gan = GanTrainer(netG, d_gpu, lr_G=1e-4, lr_D=1e-4, crop=crop, **weights)       # crop: the reference's HR_unpadder on fake_H / HR
    log = {k: float(v) for k, v in gan.step(mi.to(cuda_device), hr.to(cuda_device), interpolation=u.to(cuda_device)).items()}
    # ---- the same iteration on the CPU oracle
    w = {k: v.clone().requires_grad_(True) for k, v in wts.items()}
    fake = GCEMOracle(w, pre_pad=False, nb=nb).forward(mi)
    if crop:
        fake, hr = fake[..., crop:-crop, crop:-crop], hr[..., crop:-crop, crop:-crop]
    opt_d = torch.optim.Adam(d_cpu.parameters(), lr=1e-4)
    pr, pf = d_cpu(hr), d_cpu(fake.detach())
    interp = (u * fake.detach() + (1 - u) * hr).requires_grad_(True)
    crit = d_cpu(interp)
    gi, = torch.autograd.grad(crit, interp, torch.ones_like(crit), create_graph=True)
    gp = 10.0 * ((gi.reshape(B, -1).norm(2, dim=1) - 1) ** 2).mean()
    want = {"l_d_real": float(-2 * pr.mean()), "l_d_fake": float(2 * pf.mean()), "l_d_gp": float(gp), "D_real": float(pr.mean()), "D_fake": float(pf.mean())}
    ((-2 * pr.mean() + 2 * pf.mean()) / 2 + gp).backward()
    opt_d.step()
    for p in d_cpu.parameters():
        p.requires_grad_(False)
    l_pix, l_range = (fake - hr).abs().mean(), torch.maximum(fake - 1, -fake).clamp_min(0).mean()
    l_gan = -5e-3 * d_cpu(fake).mean()
    want.update(l_g_pix=float(l_pix), l_g_range=float(l_range), l_g_gan=float(l_gan))
    (1e-2 * l_pix + 50.0 * l_range + l_gan).backward()
    for k, v in want.items():
        assert abs(log[k] - v) <= 2e-2 * max(abs(v), 1e-3), (k, log[k], v)
    for (k, p), q in zip(d_gpu.named_parameters(), d_cpu.parameters()):           # the critic after its Adam step
        assert float((p.detach().cpu() - q).abs().max()) <= 2.5e-4, k             # lr 1e-4: one step moves every weight by <= 1e-4
    G = netG.generated_image_model
    for name, p in G.named_parameters():
        if name == "model.6.bias":
            continue
        got, ref = p.grad.cpu(), w[name].grad
        rel = float((got - ref).norm() / ref.norm().clamp_min(1e-20))
        cos = float((got * ref).sum() / (got.norm() * ref.norm()).clamp_min(1e-30))
        assert rel < 0.12 and cos > 0.992, "%s: relative error %g, cosine %g" % (name, rel, cos)
        assert float((p.detach().cpu() - wts[name]).abs().max()) > 0, name          # optimizer_G.step() reached the parameter


def test_fused_adam_updates_reach_the_kernels(cuda_device):
    """``torch.optim.Adam(fused=True)`` writes the parameters without bumping their version counters (the key of the
    packed-weight cache): the training forward re-packs every step regardless, so the next fake_H differs."""
    wts = synth.make_weights("kaiming", seed=6, nb=1)
    lr, z = synth.make_inputs(1, 16, 16, seed=6)
    netG = build_product_G(cuda_device, 1, "all_layers_HR_downscaled", wts, train=True)
    trainer = GeneratorTrainer(netG)
    opt = torch.optim.Adam(netG.generated_image_model.parameters(), lr=1e-3, fused=True)
    mi = concat_latent(lr, z).to(cuda_device)
    outs = []
    for _ in range(2):
        fake = trainer.forward(mi)
        outs.append(fake.detach().clone())
        fake.abs().mean().backward()
        trainer.backward(fake.grad)
        opt.step()
    assert float((outs[1] - outs[0]).abs().max()) > 1e-4


def test_eval_forward_after_a_fused_step_sees_the_new_weights(cuda_device):
    """After ``optimizer.step()`` of a fused optimiser (no version bump) a validation forward must use the updated
    weights: under no_grad with trainable parameters, and after ``.eval()`` with frozen ones."""
    wts = synth.make_weights("kaiming", seed=6, nb=1)
    lr, z = synth.make_inputs(1, 16, 16, seed=6)
    netG = build_product_G(cuda_device, 1, "all_layers_HR_downscaled", wts, train=True)
    G = netG.generated_image_model
    for p in G.parameters():
        p.requires_grad_(True)
    opt = torch.optim.Adam(G.parameters(), lr=1e-3, fused=True)
    mi = concat_latent(lr, z).to(cuda_device)
    out = netG(mi)
    out.abs().mean().backward()
    before = out.detach().clone()
    opt.step()
    with torch.no_grad():
        after = netG(mi)
    assert float((after - before).abs().max()) > 1e-4
    netG.eval()
    for p in G.parameters():
        p.requires_grad_(False)
    with torch.no_grad():
        ev1 = netG(mi)
    netG.train()
    with torch.no_grad():
        tr1 = netG(mi)
    assert torch.equal(tr1, after) and ev1.shape == after.shape


def test_flat_parameter_optimiser_updates_the_module(cuda_device):
    """GeneratorTrainer.flat_parameter(): every weight becomes a view of one flat fp32 parameter whose .grad is the flat
    gradient buffer; optimiser steps on it equal steps on the separate parameters, and state_dict / forward follow.
    (SGD: linear in the gradient.  Adam's first steps are +-lr * sign(g), so the fp32 summation-order noise of the chunked
    high-resolution items flips elements whose gradient is ~0.)"""
    wts = synth.make_weights("kaiming", seed=6, nb=1)
    lr, z = synth.make_inputs(1, 16, 16, seed=6)
    mi = concat_latent(lr, z).to(cuda_device)
    outs = []
    for flat in (False, True):
        netG = build_product_G(cuda_device, 1, "all_layers_HR_downscaled", wts, train=True)
        G = netG.generated_image_model
        for p in G.parameters():
            p.requires_grad_(True)
        trainer = GeneratorTrainer(netG)
        opt = torch.optim.SGD([trainer.flat_parameter()] if flat else list(G.parameters()), lr=0.05)
        for _ in range(2):
            fake = trainer.forward(mi)
            fake.abs().mean().backward()
            trainer.backward(fake.grad)
            opt.step()
        outs.append(({k: v.detach().clone() for k, v in G.state_dict().items()}, trainer.forward(mi).detach().clone()))
        trainer._state = None
    moved = 0.0
    for k in outs[0][0]:
        if "Filter" in k:
            continue
        a, b = outs[0][0][k], outs[1][0][k]
        assert float((a - b).abs().max()) <= 1e-5 * float(a.abs().max()) + 1e-7, k
        moved = max(moved, float((a.cpu() - wts[k.replace("generated_image_model.", "")]).abs().max()) if k.replace("generated_image_model.", "") in wts else 0.0)
    assert moved > 1e-5, "the optimiser steps did not reach the module's parameters"
    assert torch.allclose(outs[0][1], outs[1][1], atol=1e-4)
