"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/codes) on CPU.

TEST INFRASTRUCTURE, build-container only (the reference tree does not exist on the GPU box).
Run:  python -m oracle.gen_golden
Every fixture stores the seeds / configuration needed to regenerate the weights and inputs with
``esr_b200.synth`` plus the reference's outputs, so fixtures stay a few hundred KB in total.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402
from esr_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


make_opt, build_ref_G = ref_shims.make_opt, ref_shims.build_ref_G


def aniso_kernel(n, theta, s1, s2, shift=(0.0, 0.0)):
    """Anisotropic, off-centre Gaussian: a stand-in for an estimated (KernelGAN-style) downscaling kernel."""
    yy, xx = np.mgrid[:n, :n] - (n - 1) / 2
    xx, yy = xx - shift[0], yy - shift[1]
    a = (np.cos(theta) * xx + np.sin(theta) * yy) / s1
    b = (-np.sin(theta) * xx + np.cos(theta) * yy) / s2
    k = np.exp(-0.5 * (a * a + b * b))
    return k / k.sum()


NONDEFAULT_CASES = [  # name, sf, kernel, with operator outputs
    ("blur1_x4", 4, "blurry_cubic_1", True),        # still rank-1: separable kernels, longer taps
    ("blur2_x4", 4, "blurry_cubic_2", True),        # inv_hTh not rank-1 (magnitude clamp active)
    ("blur07_x2", 2, "blurry_cubic_0.7", False),
    ("blur1_x3", 3, "blurry_cubic_1", False),
    ("aniso13_x4", 4, aniso_kernel(13, 0.6, 3.0, 1.5), True),
    ("aniso17s_x4", 4, aniso_kernel(17, 0.6, 3.0, 1.5, (1.3, -0.8)), False),
    ("aniso15_x2", 2, aniso_kernel(15, 0.3, 2.0, 1.0, (0.4, 1.2)), True),
    ("aniso15_x3", 3, aniso_kernel(15, 1.1, 2.5, 1.2, (-1.0, 0.6)), True),
    ("aniso21s_x4", 4, aniso_kernel(21, 2.0, 4.0, 2.5, (2.2, 1.7)), False),
    # SRRaGAN_model.py:63-65: estimated kernels are inverted with lower_magnitude_bound = 0.1
    ("aniso13_x4_lmb01", 4, aniso_kernel(13, 0.6, 3.0, 1.5), False),
]


def gen_nondefault(CEMnet):
    """Non-default CEM kernels (imresize_CEM.py:22-42): derived filters, operator outputs, projection and its
    gradient, from the unmodified reference."""
    import CEM.imresize_CEM as im

    class Stub(torch.nn.Module):
        num_latent_channels, upscale = 3, 4

        def forward(self, x):
            return self.y

    out = {}
    rng = np.random.default_rng(21)
    for name, sf, kernel, with_ops in NONDEFAULT_CASES:
        im.imresize.kernels = {}
        conf = CEMnet.Get_CEM_Config(sf)
        if name.endswith("_lmb01"):
            conf.lower_magnitude_bound = 0.1
        cem = CEMnet.CEMnet(conf, upscale_kernel=kernel)
        out[name + "_sf"] = np.array(sf)
        out[name + "_lmb"] = np.array(conf.lower_magnitude_bound)
        out[name + "_kernel"] = np.array(kernel) if isinstance(kernel, str) else kernel
        out[name + "_ds_kernel"] = cem.ds_kernel.astype(np.float32)
        out[name + "_inv_hTh"] = cem.inv_hTh.astype(np.float64)
        out[name + "_margins"] = np.array([cem.invalidity_margins_LR, cem.invalidity_margins_HR,
                                           cem.ds_kernel_invalidity_half_size_LR, cem.inv_hTh_invalidity_half_size])
        print(name, cem.ds_kernel.shape, cem.inv_hTh.shape, out[name + "_margins"])
        if not with_ops:
            continue
        stub = Stub()
        wrapped = cem.WrapArchitecture_PyTorch(stub)
        h, w = 11, 14
        y = torch.from_numpy(rng.random((1, 3, sf * h, sf * w), dtype=np.float32))
        x = torch.from_numpy(rng.random((1, 3, h, w), dtype=np.float32))
        stub.y = y
        wrapped.train(True)
        with torch.no_grad():
            ops = dict(y=y.numpy(), x=x.numpy(), down=wrapped.DownscaleOP(y).numpy(), up=wrapped.Upscale_OP(x).numpy(),
                       inv=wrapped.Conv_LR_with_Inv_hTh_OP(x).numpy(), project=wrapped(x).numpy())
        yg = y.clone().requires_grad_(True)
        stub.y = yg
        g = torch.from_numpy(rng.standard_normal(tuple(y.shape)).astype(np.float32))
        (wrapped(x) * g).sum().backward()
        ops["grad_g"], ops["grad_y"] = g.numpy(), yg.grad.numpy()
        out.update({name + "_" + k: v for k, v in ops.items()})
    im.imresize.kernels = {}
    np.savez_compressed(os.path.join(OUT, "cem_nondefault.npz"), **out)


def gen_x2(CEMnet, networks):
    """x2 generator + CEM (one nearest-x2 upconv stage, architecture.py:140-146; the reference's x3 branch cannot be
    constructed): outputs in eval (pre-padded) and train mode, and the Z gradient through the reference's autograd."""
    out = {}
    for name, nb, kind, seed, B, h, w, train in (("x2_nb2_eval", 2, "default", 6, 1, 12, 10, False),
                                                 ("x2_nb1_train", 1, "kaiming", 8, 2, 9, 14, True)):
        netG, _ = build_ref_G(CEMnet, networks, nb, "all_layers", kind, seed, sf=2)
        netG.train(train)
        for p in netG.parameters():
            p.requires_grad = False
        lr, z = synth.make_inputs(B, h, w, sf=2, seed=seed)
        zg = z.clone().requires_grad_(True)
        res = netG(torch.cat([zg.contiguous().view(B, 12, h, w), lr], 1))
        g = torch.from_numpy(np.random.default_rng(seed).standard_normal(tuple(res.shape)).astype(np.float32))
        (res * g).sum().backward()
        out[name + "_out"], out[name + "_gout"], out[name + "_gz"] = res.detach().numpy(), g.numpy(), zg.grad.numpy()
        out[name + "_cfg"] = np.array([nb, seed, B, h, w, int(train)])
        out[name + "_kind"] = np.array(kind)
        print(name, tuple(res.shape), float(res.abs().max()), float(zg.grad.abs().max()))
    np.savez_compressed(os.path.join(OUT, "g_cem_x2.npz"), **out)


def gen_lr_domain(CEMnet, networks):
    """latent_input_domain 'LR' (architecture.py:137-139,159,165-166): Z at the LR image's size assigned to ``.Z``, the
    3-channel image as forward's argument.  Outputs in train and eval mode (eval: the CEM wrapper pads x by its margin, so
    Z is given at the padded size) and the Z gradient of the reference's autograd; all_layers and first_layer."""
    from oracle.ref_shims import make_opt
    import CEM.imresize_CEM as im
    out = {}
    for name, latent, nb, kind, seed, B, h, w, train in (("lr_all_nb2_train", "all_layers", 2, "default", 21, 2, 12, 10, True),
                                                        ("lr_all_nb1_eval", "all_layers", 1, "kaiming", 22, 1, 9, 14, False),
                                                        ("lr_first_nb1_train", "first_layer", 1, "default", 23, 1, 10, 12, True)):
        im.imresize.kernels = {}
        cem = CEMnet.CEMnet(CEMnet.Get_CEM_Config(4))
        opt = make_opt(nb, latent)
        opt["network_G"]["latent_input_domain"] = "LR"
        netG = networks.define_G(opt, CEM=cem, num_latent_channels=3)
        wts = synth.make_weights(kind, seed=seed, nb=nb, latent_input=latent + "_HR_downscaled")     # same conv shapes
        sd = netG.state_dict()
        assert [k for k in sd if "Filter" not in k] == ["generated_image_model." + k for k in wts]
        sd.update({"generated_image_model." + k: v for k, v in wts.items()})
        netG.load_state_dict(sd)
        netG.train(train)
        for p in netG.parameters():
            p.requires_grad = False
        m = 0 if train else int(cem.invalidity_margins_LR)
        rng = np.random.default_rng(seed)
        lr = torch.from_numpy(rng.random((B, 3, h, w), dtype=np.float32))
        z = torch.from_numpy((2 * rng.random((B, 3, h + 2 * m, w + 2 * m), dtype=np.float32) - 1))
        zg = z.clone().requires_grad_(True)
        netG.generated_image_model.Z = zg
        res = netG(lr)
        g = torch.from_numpy(rng.standard_normal(tuple(res.shape)).astype(np.float32))
        (res * g).sum().backward()
        out[name + "_lr"], out[name + "_z"] = lr.numpy(), z.numpy()
        out[name + "_out"], out[name + "_gout"], out[name + "_gz"] = res.detach().numpy(), g.numpy(), zg.grad.numpy()
        out[name + "_cfg"] = np.array([nb, seed, int(train)])
        out[name + "_kind"], out[name + "_latent"] = np.array(kind), np.array(latent)
        print(name, tuple(res.shape), float(res.abs().max()), float(zg.grad.abs().max()))
    np.savez_compressed(os.path.join(OUT, "g_cem_lr_domain.npz"), **out)


def gen_rearranged(CEMnet, networks):
    """latent_input "first_layer" with latent_input_domain "HR_rearranged" (architecture.py:109-110,159): the HR latent
    rearranged into Cz * sf^2 = 48 channels at the LR image's size, assigned to ``.Z`` and concatenated to the first conv
    only.  (all_layers_HR_rearranged raises 'Unsupported yet' in the reference's own forward.)  Train and eval outputs."""
    from oracle.ref_shims import make_opt
    import CEM.imresize_CEM as im
    out = {}
    for name, nb, kind, seed, B, h, w, train in (("rearr_nb2_train", 2, "default", 31, 2, 12, 10, True),
                                                 ("rearr_nb1_eval", 1, "kaiming", 32, 1, 9, 14, False)):
        im.imresize.kernels = {}
        cem = CEMnet.CEMnet(CEMnet.Get_CEM_Config(4))
        opt = make_opt(nb, "first_layer")
        opt["network_G"]["latent_input_domain"] = "HR_rearranged"
        netG = networks.define_G(opt, CEM=cem, num_latent_channels=3)
        wts = synth.make_weights(kind, seed=seed, nb=nb, latent_input="first_layer_HR_rearranged", num_latent_channels=48)
        sd = netG.state_dict()
        assert [k for k in sd if "Filter" not in k] == ["generated_image_model." + k for k in wts]
        assert netG.generated_image_model.num_latent_channels == 48
        sd.update({"generated_image_model." + k: v for k, v in wts.items()})
        netG.load_state_dict(sd)
        netG.train(train)
        m = 0 if train else int(cem.invalidity_margins_LR)
        rng = np.random.default_rng(seed)
        lr = torch.from_numpy(rng.random((B, 3, h, w), dtype=np.float32))
        z = torch.from_numpy((2 * rng.random((B, 48, h + 2 * m, w + 2 * m), dtype=np.float32) - 1))
        netG.generated_image_model.Z = z
        with torch.no_grad():
            res = netG(lr)
        out[name + "_lr"], out[name + "_z"], out[name + "_out"] = lr.numpy(), z.numpy(), res.numpy()
        out[name + "_cfg"] = np.array([nb, seed, int(train)])
        out[name + "_kind"] = np.array(kind)
        print(name, tuple(res.shape), float(res.abs().max()))
    np.savez_compressed(os.path.join(OUT, "g_cem_rearranged.npz"), **out)


class RefModel:
    """Minimal stand-in for SRRaGANModel (codes/models/SRRaGAN_model.py:249-302 restated): only what Z_optimizer touches."""

    def __init__(self, netG):
        self.netG, self.num_latent_channels, self.opt = netG, 3, {"scale": 4}

    def ConcatLatent(self, LR_image, latent_input):
        if LR_image.size()[2:] != latent_input.size()[2:]:
            latent_input = latent_input.contiguous().view([latent_input.size(0), latent_input.size(1) * 16] + list(LR_image.size()[2:]))
        self.model_input = torch.cat([latent_input, LR_image], dim=1)

    def GetLatent(self):
        latent = 1 * self.model_input[:, :-3, ...]
        return latent.view([latent.size(0), 3] + [4 * v for v in latent.size()[2:]])

    def feed_data(self, data, need_HR=True):
        self.var_L = data["LR"]
        self.ConcatLatent(self.var_L, data["Z"])


ZOPT2_CASES = [  # name, objective, max_iters, masks, batch
    ("max_std", "max_STD", 4, False, 1),          # Z_optimization.py:426-435, :603-607, 'max' sign flip :618
    ("min_std", "min_STD", 3, False, 1),
    ("std_increase", "STD_increase", 4, False, 1),   # additive increment data['STD_increment'] (:431-435)
    ("std_decrease_mult", "STD_decrease", 3, False, 1),   # STD_increment None: multiplicative 1/1.05 (:433)
    ("tv_converge", "TV", -3, False, 1),          # max_iters < 0: relative-decrease stop rule, cap 5*|max_iters| (:564-571)
    ("tv_masked", "TV", 4, True, 1),              # image_mask + Z_mask (:347-355; Optimizable_Z mask blend :300-303)
    ("tv_batch2", "TV", 3, False, 2),             # batch_size 2: per-image losses (latest_Z_loss_values), mean for Adam
]


def zopt2_masks(h4, w4):
    """Deterministic image / Z masks of the masked case (both HR sized, as GUI.py builds them)."""
    im = np.zeros((h4, w4), dtype=np.float32)
    im[h4 // 4:3 * h4 // 4, w4 // 8:5 * w4 // 8] = 1
    zm = np.zeros((h4, w4), dtype=np.float32)
    zm[h4 // 8:7 * h4 // 8, :3 * w4 // 4] = 1
    return im, zm


def gen_zopt2(CEMnet, networks, zopt):
    """More of the reference's Z_optimizer on the built objectives: STD objectives, convergence mode, masks, batch 2."""
    zres = {}
    for name, objective, max_iters, masked, bs in ZOPT2_CASES:
        netG, cem = build_ref_G(CEMnet, networks, 2, "all_layers", "default", 5)
        netG.train(False)
        lr, z0 = synth.make_inputs(1, 8, 8, seed=5)
        lr = lr.repeat(bs, 1, 1, 1)
        model = RefModel(netG)
        data = {"LR": lr, "Z": (0.5 * z0).repeat(bs, 1, 1, 1)}
        if "increase" in objective or "decrease" in objective:
            data["STD_increment"] = None if name.endswith("_mult") else 0.02
        model.feed_data(data)
        with torch.no_grad():
            model.fake_H = netG(model.model_input)
        kw = {}
        if masked:
            im, zm = zopt2_masks(32, 32)
            kw = dict(image_mask=im, Z_mask=zm, initial_Z=0.5 * z0)
        opt = zopt.Z_optimizer(objective=objective, Z_size=[32, 32], model=model, Z_range=1.0, max_iters=max_iters, data=data,
                               initial_LR=0.1, batch_size=bs, **kw)
        if bs > 1:       # pin the per-image initial Z (the reference would draw it with torch's RNG, :559-560)
            opt.random_Z_inits = False
            z_init = torch.from_numpy(np.random.default_rng(17).standard_normal((bs, 3, 32, 32)).astype(np.float32)) * 0.3
            opt.Z_model.Z.data.copy_(z_init)
            zres[name + "_Zinit"] = z_init.numpy()
        Z = opt.optimize()
        zres[name + "_loss"] = np.array(opt.loss_values, dtype=np.float64)
        zres[name + "_latest"] = np.array(opt.latest_Z_loss_values, dtype=np.float64)
        zres[name + "_Z"] = Z.numpy()
        zres[name + "_cur_iter"] = np.array(opt.cur_iter)
        zres[name + "_initial_STD"] = opt.initial_STD.numpy()
        print(name, opt.loss_values, opt.cur_iter)
    np.savez_compressed(os.path.join(OUT, "zopt2.npz"), **zres)


def gen_zobj(CEMnet, networks, zopt):
    """The rich objectives (Z_optimization.py:21-270, :371-523, :576-617) from the unmodified reference: SoftHistogramLoss
    values / gradients / bins on the HIST_CASES of oracle/zobj_cases.py, and Z_optimizer trajectories of the ZOPT3 cases
    through the reference's nb=2 G+CEM (GUI mode).  'scribble' has no reference run (zobj_cases.NO_REFERENCE_RUN)."""
    from oracle import zobj_cases as zc
    res = {}
    for name in zc.HIST_CASES:
        value, grad, bins = zc.run_hist_case(zopt.SoftHistogramLoss, name)
        res["hist_%s_value" % name], res["hist_%s_grad" % name] = value.numpy().astype(np.float64), grad.numpy()
        res["hist_%s_bins" % name] = bins.numpy()
        print("hist", name, value, float(grad.abs().max()), tuple(bins.shape))
    lr, z0 = synth.make_inputs(1, zc.ZOPT3_HW[0], zc.ZOPT3_HW[1], seed=11)
    for name in zc.ZOPT3_CASES:
        if name in zc.NO_REFERENCE_RUN:
            continue
        netG, cem = build_ref_G(CEMnet, networks, 2, "all_layers", "default", 5)
        netG.train(False)
        opt, Z = zc.run_zopt_case(zopt.Z_optimizer, RefModel(netG), netG, name, lr, z0, z_init=zc.zopt3_z_init(name))
        res["zopt_%s_loss" % name] = np.array(opt.loss_values, dtype=np.float64)
        res["zopt_%s_latest" % name] = np.array(opt.latest_Z_loss_values, dtype=np.float64).reshape(-1)
        res["zopt_%s_Z" % name] = Z.numpy()
        res["zopt_%s_initial_STD" % name] = opt.initial_STD.numpy()
        print("zopt", name, opt.loss_values)
    np.savez_compressed(os.path.join(OUT, "zobjectives.npz"), **res)


CFG3_WINDOWS = [(0, 0), (464, 464), (928, 928), (0, 928)]      # top-left corners of the stored 96 x 96 HR windows
CFG3_WIN = 96


def gen_cfg3(CEMnet, networks):
    """BASELINE config 3 at its own size: the production generator (nb = 23, default init) + CEM, eval mode, one
    1x3x256x256 LR image; output and dL/dZ of the reference's autograd for L = sum(out * g).  Only windows of the two
    1x3x1024x1024 tensors are stored (borders, centre) plus whole-tensor norms."""
    netG, _ = build_ref_G(CEMnet, networks, 23, "all_layers", "default", 0)
    netG.train(False)
    for p in netG.parameters():
        p.requires_grad = False
    lr, z = synth.make_inputs(1, 256, 256, seed=33)
    zg = z.clone().requires_grad_(True)
    res = netG(torch.cat([zg.contiguous().view(1, 48, 256, 256), lr], 1))
    g = torch.from_numpy(np.random.default_rng(33).standard_normal(tuple(res.shape)).astype(np.float32))
    (res * g).sum().backward()
    out = {"cfg": np.array([23, 0, 33, 256, 256]), "out_norm": np.array(float(res.detach().double().norm())),
           "gz_norm": np.array(float(zg.grad.double().norm())), "out_mean": np.array(float(res.detach().double().mean()))}
    for k, (y0, x0) in enumerate(CFG3_WINDOWS):
        out["out_%d" % k] = res.detach()[0, :, y0:y0 + CFG3_WIN, x0:x0 + CFG3_WIN].numpy()
        out["gz_%d" % k] = zg.grad[0, :, y0:y0 + CFG3_WIN, x0:x0 + CFG3_WIN].numpy()
    print("cfg3", tuple(res.shape), out["out_norm"], out["gz_norm"])
    np.savez_compressed(os.path.join(OUT, "cfg3.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    CEMnet, networks, arch, zopt = ref_shims.load_reference()
    import CEM.imresize_CEM as im

    if "nondefault" in sys.argv[1:]:      # only the non-default-kernel fixture
        return gen_nondefault(CEMnet)
    if "x2" in sys.argv[1:]:              # only the x2 fixture
        return gen_x2(CEMnet, networks)
    if "rearranged" in sys.argv[1:]:      # only the first_layer_HR_rearranged fixture
        return gen_rearranged(CEMnet, networks)
    if "lr_domain" in sys.argv[1:]:       # only the LR-domain latent fixture
        return gen_lr_domain(CEMnet, networks)
    if "zopt2" in sys.argv[1:]:           # only the extra Z_optimizer trajectories
        return gen_zopt2(CEMnet, networks, zopt)
    if "zobj" in sys.argv[1:]:            # only the rich-objective fixture
        return gen_zobj(CEMnet, networks, zopt)
    if "cfg3" in sys.argv[1:]:            # only the config-3-size output / gradient windows (about a minute, ~20 GB)
        return gen_cfg3(CEMnet, networks)

    # 1. filters --------------------------------------------------------------------------
    filt = {}
    for sf in (2, 3, 4):
        im.imresize.kernels = {}
        c = CEMnet.CEMnet(CEMnet.Get_CEM_Config(sf))
        filt["ds_kernel_%d" % sf] = c.ds_kernel.astype(np.float32)
        filt["inv_hTh_%d" % sf] = c.inv_hTh.astype(np.float64)
        filt["margins_%d" % sf] = np.array([c.invalidity_margins_LR, c.invalidity_margins_HR,
                                            c.ds_kernel_invalidity_half_size_LR, c.inv_hTh_invalidity_half_size])
    np.savez_compressed(os.path.join(OUT, "cem_filters.npz"), **filt)

    # 2. CEM operators and projection around a stub generator -----------------------------
    class Stub(torch.nn.Module):
        num_latent_channels, upscale = 3, 4

        def forward(self, x):
            return self.y

    im.imresize.kernels = {}
    cem = CEMnet.CEMnet(CEMnet.Get_CEM_Config(4))
    stub = Stub()
    wrapped = cem.WrapArchitecture_PyTorch(stub)
    rng = np.random.default_rng(7)
    y = torch.from_numpy(rng.random((2, 3, 48, 64), dtype=np.float32))
    x = torch.from_numpy(rng.random((2, 3, 12, 16), dtype=np.float32))
    stub.y = y
    wrapped.train(True)
    with torch.no_grad():
        ops = dict(y=y.numpy(), x=x.numpy(), down=wrapped.DownscaleOP(y).numpy(), up=wrapped.Upscale_OP(x).numpy(),
                   inv=wrapped.Conv_LR_with_Inv_hTh_OP(x).numpy(), project_train=wrapped(x).numpy())
    # adjoint of the projection w.r.t. y through autograd
    yg = y.clone().requires_grad_(True)
    stub.y = yg
    g = torch.from_numpy(rng.standard_normal((2, 3, 48, 64)).astype(np.float32))
    (wrapped(x) * g).sum().backward()
    ops["project_grad_g"], ops["project_grad_y"] = g.numpy(), yg.grad.numpy()
    # eval mode: padded by 10/40, Z packed in front
    lr, z = synth.make_inputs(1, 8, 12, seed=3)
    mi = torch.cat([z.contiguous().view(1, 48, 8, 12), lr], 1)
    stub.y = torch.from_numpy(rng.random((1, 3, 4 * 28, 4 * 32), dtype=np.float32))
    wrapped.train(False)
    with torch.no_grad():
        ops["eval_model_input"], ops["eval_y"], ops["project_eval"] = mi.numpy(), stub.y.numpy(), wrapped(mi).numpy()
    np.savez_compressed(os.path.join(OUT, "cem_ops.npz"), **ops)

    # 3. G + CEM --------------------------------------------------------------------------
    cases = [("prod_default", 23, "all_layers", "default", 0, 1, 12, 12, False),
             ("prod_kaiming", 23, "all_layers", "kaiming", 1, 1, 10, 14, False),
             ("nb2_train_mode", 2, "all_layers", "default", 2, 2, 16, 12, True),
             ("nb2_first_layer", 2, "first_layer", "default", 3, 1, 12, 12, False),
             ("nb1_no_latent", 1, "None", "default", 4, 1, 12, 16, False)]
    out = {}
    for name, nb, li, kind, seed, B, h, w, train in cases:
        netG, _ = build_ref_G(CEMnet, networks, nb, li, kind, seed)
        netG.train(train)
        lr, z = synth.make_inputs(B, h, w, seed=seed)
        mi = lr if li == "None" else torch.cat([z.contiguous().view(B, 48, h, w), lr], 1)
        with torch.no_grad():
            res = netG(mi)
        out[name + "_out"] = res.numpy()
        out[name + "_cfg"] = np.array([nb, seed, B, h, w, int(train)])
        out[name + "_str"] = np.array([li, kind])
        print(name, tuple(res.shape), float(res.abs().max()))
    np.savez_compressed(os.path.join(OUT, "g_cem.npz"), **out)

    # 4. Z optimisation through the reference's Z_optimizer --------------------------------
    zres = {}
    for name, objective, train_mode in (("tv_eval", "TV", False), ("l1_train", "l1", True)):
        netG, cem = build_ref_G(CEMnet, networks, 2, "all_layers", "default", 5)
        netG.train(train_mode)
        lr, z0 = synth.make_inputs(1, 8, 8, seed=5)
        model = RefModel(netG)
        data = {"LR": lr, "Z": 0.5 * z0}
        model.feed_data(data)
        with torch.no_grad():
            model.fake_H = netG(model.model_input)
        if train_mode:
            del model.__dict__["fake_H"]
            tgt = torch.from_numpy(np.random.default_rng(11).random((1, 3, 32, 32), dtype=np.float32))
            data["HR"] = tgt
            zres[name + "_target"] = tgt.numpy()
        opt = zopt.Z_optimizer(objective=objective, Z_size=[32, 32], model=model, Z_range=1.0, max_iters=4, data=data,
                               initial_LR=0.1, batch_size=1, HR_unpadder=(lambda t: t) if train_mode else None)
        if train_mode:
            opt.feed_data(data)
            # training mode re-draws Z with torch's RNG (Z_optimization.py:559-560); pin it to a numpy draw
            opt.random_Z_inits = False
            z_init = torch.from_numpy(np.random.default_rng(13).standard_normal((1, 3, 32, 32)).astype(np.float32))
            opt.Z_model.Z.data.copy_(z_init)
            zres[name + "_Zinit"] = z_init.numpy()
        Z = opt.optimize()
        zres[name + "_loss"] = np.array(opt.loss_values, dtype=np.float64)
        zres[name + "_Z"] = Z.numpy()
        print(name, opt.loss_values)
    np.savez_compressed(os.path.join(OUT, "zopt.npz"), **zres)

    # 5. non-default CEM kernels -----------------------------------------------------------
    gen_nondefault(CEMnet)

    # 6. x2 generator ------------------------------------------------------------------------
    gen_x2(CEMnet, networks)

    # 7. more Z_optimizer paths, config-3-size gradient ----------------------------------------
    gen_zopt2(CEMnet, networks, zopt)
    gen_cfg3(CEMnet, networks)

    # 8. LR-domain latent input ------------------------------------------------------------------
    gen_lr_domain(CEMnet, networks)
    gen_rearranged(CEMnet, networks)
    gen_zobj(CEMnet, networks, zopt)


if __name__ == "__main__":
    main()
