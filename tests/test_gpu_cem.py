"""CEM CUDA operators against the reference's golden outputs and the CPU oracle (fp32, <=1e-5)."""
import numpy as np
import pytest
import torch

from esr_b200 import cem as pcem
from oracle.cem_ops import CEMOracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wrapped(cuda_device):
    class Stub(torch.nn.Module):
        num_latent_channels, upscale = 3, 4

        def forward(self, x):
            return self.y
    net = pcem.CEMnet(pcem.Get_CEM_Config(4))
    stub = Stub()
    w = net.WrapArchitecture_PyTorch(stub).to(cuda_device)
    return w, stub, net


def test_ops_match_reference_golden(golden, wrapped, cuda_device):
    w, stub, _ = wrapped
    g = golden("cem_ops")
    y, x = torch.from_numpy(g["y"]).to(cuda_device), torch.from_numpy(g["x"]).to(cuda_device)
    with torch.no_grad():
        np.testing.assert_allclose(w.DownscaleOP(y).cpu().numpy(), g["down"], atol=1e-5)
        np.testing.assert_allclose(w.Upscale_OP(x).cpu().numpy(), g["up"], atol=1e-5)
        np.testing.assert_allclose(w.Conv_LR_with_Inv_hTh_OP(x).cpu().numpy(), g["inv"], atol=1e-5)
        stub.y = y
        w.train(True)
        np.testing.assert_allclose(w(x).cpu().numpy(), g["project_train"], atol=1e-5)
        stub.y = torch.from_numpy(g["eval_y"]).to(cuda_device)
        w.train(False)
        out = w(torch.from_numpy(g["eval_model_input"]).to(cuda_device))
        np.testing.assert_allclose(out.cpu().numpy(), g["project_eval"], atol=1e-5)


def test_projection_gradient_matches_reference_autograd(golden, wrapped, cuda_device):
    w, stub, _ = wrapped
    g = golden("cem_ops")
    y = torch.from_numpy(g["y"]).to(cuda_device).requires_grad_(True)
    stub.y = y
    w.train(True)
    out = w(torch.from_numpy(g["x"]).to(cuda_device))
    (out * torch.from_numpy(g["project_grad_g"]).to(cuda_device)).sum().backward()
    np.testing.assert_allclose(y.grad.cpu().numpy(), g["project_grad_y"], atol=1e-5)


@pytest.mark.parametrize("sf,shape", [(4, (1, 3, 2048, 2048)), (4, (2, 3, 100, 36)), (2, (1, 3, 64, 96))])
def test_ops_match_oracle_and_consistency(cuda_device, sf, shape):
    """BASELINE config 4 size included; consistency ||Down(x_hat) - x||_inf <= 1e-4 on the interior."""
    net = pcem.CEMnet(pcem.Get_CEM_Config(sf))

    class Stub(torch.nn.Module):
        num_latent_channels, upscale = 3, sf

        def forward(self, x):
            return self.y
    stub = Stub()
    w = net.WrapArchitecture_PyTorch(stub).to(cuda_device).train(True)
    gen = torch.Generator().manual_seed(0)
    B, Cc, H, W = shape
    y = torch.rand(shape, generator=gen)
    x = torch.rand(B, Cc, H // sf, W // sf, generator=gen)
    stub.y = y.to(cuda_device)
    with torch.no_grad():
        out = w(x.to(cuda_device))
        res = (w.DownscaleOP(out) - x.to(cuda_device)).abs()
    m = 3 if sf == 4 else 4          # the reference's own residual reaches 6e-4 at 3 LR px for x2 (oracle agrees)
    assert res[:, :, m:-m, m:-m].max().item() <= 1e-4
    # against the CPU oracle at every size, BASELINE config 4 (1x3x2048x2048) included: ~0.5 s of CPU there
    ora = CEMOracle(sf)
    np.testing.assert_allclose(out.cpu().numpy(), ora.project(y, x).numpy(), atol=1e-5)
    if H > 256:  # and a size-independent property at full size: linearity of the projection in (y, x)
        y2 = torch.rand(shape, generator=gen).to(cuda_device)
        with torch.no_grad():
            stub.y = y2
            out2 = w(torch.zeros_like(x).to(cuda_device))
            stub.y = stub.y + y.to(cuda_device)
            out12 = w(x.to(cuda_device))
        assert (out12 - out - out2).abs().max().item() <= 2e-5


@pytest.mark.parametrize("shape,crop", [((1, 3, 2048, 2048), 0),      # BASELINE config 4: 3 rounds of one plane
                                        ((2, 3, 1024, 1088), 0),       # two planes per round; W not a multiple of 256
                                        ((1, 3, 112, 1344), 40),       # eval-mode crop, 7 slabs per plane, all planes in one round
                                        ((1, 3, 16, 1024), 0),         # a single slab: top and bottom replicate rows in the same CTA
                                        ((2, 3, 2368, 1024), 8)])      # 148 slabs per plane: every SM of the part
def test_single_launch_projection(cuda_device, shape, crop):
    """esr_cem_project_fused (csrc/cem_fused.cu: one cooperative launch, y read once, u = K_h(x - Down y) accumulated
    through L2) against the CPU oracle (1e-5) and the default two-launch path; run-to-run identical."""
    from esr_b200 import _capi as capi
    B, Cc, H, W = shape
    gen = torch.Generator().manual_seed(5)
    y = torch.rand(shape, generator=gen)
    x = torch.rand(B, Cc, H // 4, W // 4, generator=gen)
    net = pcem.CEMnet(pcem.Get_CEM_Config(4))
    f = net._filters
    yd, xd = y.to(cuda_device), x.to(cuda_device)
    out = torch.full((B, Cc, H - 2 * crop, W - 2 * crop), float("nan"), device=cuda_device)
    ws = torch.full((2 * B * Cc * (H // 4) * (W // 4),), float("nan"), device=cuda_device)       # the launch must not rely on a clean workspace
    with torch.cuda.device(cuda_device):
        for _ in range(2):                                                                            # and must leave it reusable
            capi.cem_call("project_fused", f, capi.ptr(yd), capi.ptr(xd), B, Cc, H, W, crop, capi.ptr(out), capi.ptr(ws), capi.stream_ptr())
    torch.cuda.synchronize()
    ref = CEMOracle(4).project(y, x)
    if crop:
        ref = ref[..., crop:H - crop, crop:W - crop]
    got = out.cpu()
    assert torch.isfinite(got).all()
    np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=1e-5)
    # run-to-run identical (two order-independent contributions per cell, no other atomics)
    out2, out3 = torch.empty_like(out), torch.empty_like(out)
    with torch.cuda.device(cuda_device):
        capi.cem_call("project_fused", f, capi.ptr(yd), capi.ptr(xd), B, Cc, H, W, crop, capi.ptr(out2), capi.ptr(ws), capi.stream_ptr())
        capi.cem_call("project", f, capi.ptr(yd), capi.ptr(xd), B, Cc, H, W, crop, capi.ptr(out3), capi.ptr(ws), capi.stream_ptr())
    assert torch.equal(out, out2)
    assert float((out - out3).abs().max()) <= 1e-5
    with pytest.raises(capi.EsrError, match="single-launch"):      # a shape outside its domain is refused, not mangled
        capi.cem_call("project_fused", f, capi.ptr(yd), capi.ptr(xd), B, Cc, H // 2 + 4, W, 0, capi.ptr(out2), capi.ptr(ws), capi.stream_ptr())


# ------------------------------------------------------------------ non-default kernels (csrc/cem2d.cu)
NONDEFAULT_OP_CASES = ["blur1_x4", "blur2_x4", "aniso13_x4", "aniso15_x2", "aniso15_x3"]


def _kernel_of(g, name):
    k = g[name + "_kernel"]
    return str(k) if k.dtype.kind in "US" else k


def _stub_wrapped(net, sf, dev, train=True):
    class Stub(torch.nn.Module):
        num_latent_channels, upscale = 3, sf

        def forward(self, x):
            return self.y
    stub = Stub()
    return net.WrapArchitecture_PyTorch(stub).to(dev).train(train), stub


def _tol(net):
    """fp32 accumulation through K = inv_hTh: rounding of the O(1) operands is amplified by ||K||_1 (15 for the
    mild kernels, 220 for blurry_cubic_2, whose inverse is ill conditioned by construction)."""
    return 1e-5 + 3e-7 * float(np.abs(net.inv_hTh).sum())


def _oracle_of(net):
    sf = int(net.ds_factor)
    return CEMOracle(sf, filters=dict(ds_kernel=net.ds_kernel, inv_hTh=net.inv_hTh, margin_LR=int(net.invalidity_margins_LR),
                                      margin_HR=int(net.invalidity_margins_HR)))


@pytest.mark.parametrize("name", NONDEFAULT_OP_CASES)
def test_nondefault_ops_match_reference_golden(golden, cuda_device, name):
    """blurry_cubic_<sigma> and ndarray kernels against the unmodified reference (ops, projection, its gradient)."""
    g = golden("cem_nondefault")
    sf = int(g[name + "_sf"])
    net = pcem.CEMnet(pcem.Get_CEM_Config(sf), upscale_kernel=_kernel_of(g, name))
    w, stub = _stub_wrapped(net, sf, cuda_device)
    y, x = torch.from_numpy(g[name + "_y"]).to(cuda_device), torch.from_numpy(g[name + "_x"]).to(cuda_device)
    with torch.no_grad():
        np.testing.assert_allclose(w.DownscaleOP(y).cpu().numpy(), g[name + "_down"], atol=1e-5)
        np.testing.assert_allclose(w.Upscale_OP(x).cpu().numpy(), g[name + "_up"], atol=1e-5)
        np.testing.assert_allclose(w.Conv_LR_with_Inv_hTh_OP(x).cpu().numpy(), g[name + "_inv"], atol=_tol(net))
        stub.y = y
        np.testing.assert_allclose(w(x).cpu().numpy(), g[name + "_project"], atol=_tol(net))
    yg = y.clone().requires_grad_(True)
    stub.y = yg
    (w(x) * torch.from_numpy(g[name + "_grad_g"]).to(cuda_device)).sum().backward()
    np.testing.assert_allclose(yg.grad.cpu().numpy(), g[name + "_grad_y"], atol=_tol(net))


@pytest.mark.parametrize("name,shape,crop", [("aniso17s_x4", (2, 3, 37, 45), 0), ("aniso17s_x4", (1, 3, 70, 33), 8),
                                             ("blur2_x4", (1, 2, 40, 77), 12), ("aniso15_x3", (1, 3, 35, 67), 6),
                                             ("aniso15_x2", (3, 1, 64, 31), 2), ("aniso21s_x4", (1, 3, 5, 3), 0)])
def test_nondefault_project_and_adjoint_match_oracle(golden, cuda_device, name, shape, crop):
    """Ragged sizes (tiles cut by the image edge, images narrower than the stencils), cropped output, and the
    exact adjoint including the replicate-padding folds, against the oracle's autograd."""
    g = golden("cem_nondefault")
    sf = int(g[name + "_sf"])
    net = pcem.CEMnet(pcem.Get_CEM_Config(sf), upscale_kernel=_kernel_of(g, name))
    assert not net.separable
    ora = _oracle_of(net)
    gen = torch.Generator().manual_seed(3)
    B, Cc, h, w_ = shape
    y = torch.rand(B, Cc, sf * h, sf * w_, generator=gen)
    x = torch.rand(B, Cc, h, w_, generator=gen)

    def tile3(k):  # the oracle's depthwise weights are built for 3 planes; rebuild for Cc
        return k[:1].repeat(Cc, 1, 1, 1)
    ora.w_inv, ora.w_down, ora.w_up = tile3(ora.w_inv), tile3(ora.w_down), tile3(ora.w_up)
    ora._dw = lambda t, wt, pad: torch.nn.functional.conv2d(torch.nn.functional.pad(t, (pad,) * 4, mode="replicate"), wt, groups=Cc)
    yo = y.clone().requires_grad_(True)
    ref = ora.project(yo, x)
    ref = ref[:, :, crop:ref.size(2) - crop, crop:ref.size(3) - crop]
    gout = torch.randn(ref.shape, generator=gen)
    (ref * gout).sum().backward()
    yd = y.to(cuda_device).requires_grad_(True)
    out = pcem._CemProject.apply(yd, x.to(cuda_device), net._filters, crop)
    (out * gout.to(cuda_device)).sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), atol=_tol(net))
    np.testing.assert_allclose(yd.grad.cpu().numpy(), yo.grad.numpy(), atol=_tol(net) * max(1.0, yo.grad.abs().max().item() / 4))


def test_stencil_path_equals_separable_path(cuda_device):
    """The default bicubic filters pushed through the general 2-D stencils give what the rank-1 kernels give
    (forward, every operator, and the adjoint), at a size with many tiles per plane."""
    from esr_b200 import _capi as capi
    net = pcem.CEMnet(pcem.Get_CEM_Config(4))
    bank = capi.CemFilterBank2D(4, net.pre_stride, net.ds_kernel, net.inv_hTh)
    gen = torch.Generator().manual_seed(5)
    y = torch.rand(2, 3, 512, 384, generator=gen).to(cuda_device)
    x = torch.rand(2, 3, 128, 96, generator=gen).to(cuda_device)
    gout = torch.randn(2, 3, 512 - 80, 384 - 80, generator=gen).to(cuda_device)
    res = []
    for f in (net._filters, bank):
        yd = y.clone().requires_grad_(True)
        out = pcem._CemProject.apply(yd, x, f, 40)
        (out * gout).sum().backward()
        ops = [pcem.Filter_Layer(net.inv_hTh, op, f).to(cuda_device)(t) for op, t in (("down", y), ("up", x), ("inv", x))]
        res.append([out.detach(), yd.grad] + ops)
    for a, b in zip(*res):
        assert (a - b).abs().max().item() <= 2e-5 * max(1.0, a.abs().max().item())


def test_generator_with_estimated_kernel_matches_oracle(golden, cuda_device):
    """G + CEM built on a user-supplied kernel, eval mode (pre-padding by that kernel's own margins), forward and
    the Z gradient: the whole drop-in path with a non-default CEM."""
    from esr_b200 import networks, synth
    from oracle.rrdbnet import GCEMOracle
    from oracle.cem_ops import concat_latent
    g = golden("cem_nondefault")
    net = pcem.CEMnet(pcem.Get_CEM_Config(4), upscale_kernel=g["aniso13_x4_kernel"])
    opt = {"gpu_ids": None, "is_train": False, "datasets": {"train": {"patch_size": 256}},
           "network_G": dict(which_model_G="RRDB_net", CEM_arch=1, latent_input="all_layers", latent_input_domain="HR_downscaled",
                             latent_channels=3, norm_type=None, mode="CNA", nf=64, nb=1, in_nc=3, out_nc=3, gc=32, scale=4)}
    netG = networks.define_G(opt, CEM=net, num_latent_channels=3).to(cuda_device).eval()
    wts = synth.make_weights("default", seed=9, nb=1, latent_input="all_layers_HR_downscaled")
    sd = netG.state_dict()
    sd.update({"generated_image_model." + k: v for k, v in wts.items()})
    netG.load_state_dict(sd)
    for p in netG.parameters():
        p.requires_grad_(False)
    lr, z = synth.make_inputs(1, 9, 11, seed=9)
    ora = GCEMOracle(wts, nb=1, cem=_oracle_of(net))
    zo = z.clone().requires_grad_(True)
    ref = ora.forward(concat_latent(lr, zo))
    gout = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1))
    (ref * gout).sum().backward()
    zd = z.to(cuda_device).requires_grad_(True)
    out = netG(concat_latent(lr.to(cuda_device), zd))
    assert out.shape == ref.shape
    (out * gout.to(cuda_device)).sum().backward()
    assert (out.detach().cpu() - ref.detach()).abs().max().item() <= 1e-2
    gz, gr = zd.grad.cpu(), zo.grad
    assert float((gz - gr).norm() / gr.norm()) < 4e-2 and float((gz * gr).sum() / (gz.norm() * gr.norm())) > 0.999


def _estimated_kernel_G(golden, dev, nb=1, seed=9):
    from esr_b200 import networks, synth
    g = golden("cem_nondefault")
    conf = pcem.Get_CEM_Config(4)
    conf.lower_magnitude_bound = 0.1                     # what the reference sets for estimated kernels
    net = pcem.CEMnet(conf, upscale_kernel=g["aniso13_x4_kernel"])
    opt = {"gpu_ids": None, "is_train": False, "datasets": {"train": {"patch_size": 256}},
           "network_G": dict(which_model_G="RRDB_net", CEM_arch=1, latent_input="all_layers", latent_input_domain="HR_downscaled",
                             latent_channels=3, norm_type=None, mode="CNA", nf=64, nb=nb, in_nc=3, out_nc=3, gc=32, scale=4)}
    netG = networks.define_G(opt, CEM=net, num_latent_channels=3).to(dev).eval()
    wts = synth.make_weights("default", seed=seed, nb=nb, latent_input="all_layers_HR_downscaled")
    sd = netG.state_dict()
    sd.update({"generated_image_model." + k: v for k, v in wts.items()})
    netG.load_state_dict(sd)
    for p in netG.parameters():
        p.requires_grad_(False)
    return netG, net


def test_z_optimizer_with_estimated_kernel(golden, cuda_device):
    """The Z-optimisation loop through a CEM built on an estimated kernel: the captured graphs (which then contain the
    2-D stencil kernels and their adjoint) reproduce the eager loop's loss trajectory, and the loss goes down."""
    from esr_b200 import synth
    from esr_b200.z_optimization import Z_optimizer, SRModelShim
    losses = {}
    for mode in ("eager", "graph"):
        netG, net = _estimated_kernel_G(golden, cuda_device)
        assert not net.separable
        lr, z0 = synth.make_inputs(1, 8, 9, seed=5)
        model = SRModelShim(netG)
        data = {"LR": lr.to(cuda_device), "Z": (0.5 * z0).to(cuda_device)}
        model.feed_data(data)
        with torch.no_grad():
            model.fake_H = netG(model.model_input)
        opt = Z_optimizer(objective="TV", Z_size=[32, 36], model=model, Z_range=1.0, max_iters=4, data=data, initial_LR=0.1,
                          batch_size=1)
        if mode == "eager":
            import os
            os.environ["ESR_ZOPT_GRAPH"] = "0"
        try:
            opt.optimize()
        finally:
            if mode == "eager":
                del os.environ["ESR_ZOPT_GRAPH"]
        assert netG.generated_image_model.use_cuda_graphs == (mode == "graph")
        losses[mode] = np.array(opt.loss_values)
    np.testing.assert_allclose(losses["graph"], losses["eager"], rtol=1e-4)
    assert losses["graph"][-1] < losses["graph"][0]


def test_host_pipeline_with_estimated_kernel(golden, cuda_device):
    """Serving path (captured inference graph, copy streams) with 2-D CEM filters equals the direct call."""
    from esr_b200 import synth
    from esr_b200.parallel import HostPipeline
    from oracle.cem_ops import concat_latent
    netG, net = _estimated_kernel_G(golden, cuda_device)
    lr, z = synth.make_inputs(3, 12, 10, seed=2)
    mi = concat_latent(lr, z).contiguous()
    with torch.no_grad():
        ref = netG(mi.to(cuda_device)).cpu()
    host_in, host_out = mi.pin_memory(), torch.empty_like(ref).pin_memory()
    pipe = HostPipeline(netG, chunk=2)
    pipe(host_in, host_out)
    pipe.wait()
    assert torch.equal(host_out, ref)


def test_projection_config2_shape_many_ctas_per_sm(cuda_device):
    """16 x 3 planes of 592^2 cropped by 40 (the CEM call inside BASELINE config 2): 864 work items on co-resident CTAs.
    Regression test of a shared-memory ring race in the TMA streaming kernels (a refill landing under loads that had
    been issued but not performed: sporadic wrong rows, only with several CTAs per SM) - three runs, each against the
    CPU oracle through the C ABI."""
    import ctypes as C
    from esr_b200 import _capi as capi
    f = pcem.CEMnet(pcem.Get_CEM_Config(4))._filters
    B, Cc, H, W, crop = 16, 3, 592, 592, 40
    g = torch.Generator().manual_seed(5)
    y, x = torch.rand(B, Cc, H, W, generator=g), torch.rand(B, Cc, H // 4, W // 4, generator=g)
    ref = CEMOracle(4).project(y, x)[:, :, crop:-crop, crop:-crop]
    yd, xd = y.to(cuda_device), x.to(cuda_device)
    ws = torch.empty(2 * B * Cc * (H // 4) * (W // 4), device=cuda_device)
    for _ in range(3):
        out = torch.full((B, Cc, H - 2 * crop, W - 2 * crop), float("nan"), device=cuda_device)
        capi.check(capi.lib().esr_cem_project(f, capi.ptr(yd), capi.ptr(xd), B, Cc, H, W, crop, capi.ptr(out), capi.ptr(ws),
                                              capi.stream_ptr()))
        torch.cuda.synchronize()
        rec = (C.c_uint32 * 4)()
        capi.check(capi.lib().esr_debug_cem_timeout(rec))
        assert list(rec) == [0, 0, 0, 0], "ring wait timed out: %s" % list(rec)
        np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), atol=1e-5)
