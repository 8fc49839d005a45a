"""Full RRDBNet(+Z) -> CEM forward on the GPU against (a) the reference's golden outputs, (b) the fp32
CPU oracle and (c) the oracle emulating bf16 MMA operands (tight: separates indexing from rounding).

Tolerances (BASELINE.json north_star): PSNR >= 50 dB and max|err| <= 1e-2 against the fp32 reference on
[0,1]-scaled images; CEM consistency residual <= 1e-4 on the interior (>= 3 LR px from the border)."""
import numpy as np
import pytest
import torch

from esr_b200 import cem as pcem, networks, synth
from oracle.cem_ops import concat_latent
from oracle.rrdbnet import GCEMOracle
from tests.helpers import psnr
from tests.test_oracle_golden import CASES, oracle_for_case

pytestmark = pytest.mark.gpu


def build_product_G(dev, nb, latent, weights, train=False, sf=4, domain="HR_downscaled"):
    opt = {"gpu_ids": None, "is_train": False, "datasets": {"train": {"patch_size": 256}},
           "network_G": dict(which_model_G="RRDB_net", CEM_arch=1,
                             latent_input="None" if latent is None else latent.split("_HR_")[0],
                             latent_input_domain=domain, latent_channels=3, norm_type=None, mode="CNA",
                             nf=64, nb=nb, in_nc=3, out_nc=3, gc=32, scale=sf)}
    cemnet = pcem.CEMnet(pcem.Get_CEM_Config(sf))
    netG = networks.define_G(opt, CEM=cemnet, num_latent_channels=3 if latent else 0)
    sd = netG.state_dict()
    keys = [k for k in sd if "Filter" not in k]
    assert keys == ["generated_image_model." + k for k in weights], "state_dict key order differs from the reference"
    sd.update({"generated_image_model." + k: v for k, v in weights.items()})
    netG.load_state_dict(sd)
    netG.to(dev)
    netG.train(train)
    for p in netG.parameters():
        p.requires_grad_(False)
    return netG


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_forward_matches_reference_golden(golden, cuda_device, name, impl):
    g = golden("g_cem")
    ora, mi, wts, cfg = oracle_for_case(g, name)
    if impl == "simt" and cfg["nb"] > 2:
        pytest.skip("SIMT cross-check only on the small nets")
    netG = build_product_G(cuda_device, cfg["nb"], cfg["latent"], wts, train=cfg["train"])
    netG.generated_image_model.debug_simt = impl == "simt"
    with torch.no_grad():
        out = netG(mi.to(cuda_device)).cpu()
    ref = torch.from_numpy(g[name + "_out"])
    err, p = (out - ref).abs().max().item(), psnr(out, ref)
    assert err <= 1e-2 and p >= 50.0, "max|err| %g, PSNR %.1f dB vs the reference" % (err, p)
    # tight check against the oracle emulating the kernels' arithmetic (bf16 trunk operands; the six
    # outer convs run with fp16 operands)
    ora_b, _, _, _ = oracle_for_case(g, name, operand_dtype=torch.bfloat16)
    orig_conv = ora_b.net.conv
    outer_keys = {"model.0", "model.1.sub.%d" % cfg["nb"], "model.2.1", "model.3.1", "model.4", "model.6"}

    def conv(x, key, act):
        ora_b.net.od = torch.float16 if key in outer_keys else torch.bfloat16
        return orig_conv(x, key, act)
    ora_b.net.conv = conv
    with torch.no_grad():
        emu = ora_b.forward(mi)
    err_e = (out - emu).abs().max().item()
    assert err_e <= 2e-3, "max|err| %g vs bf16-emulating oracle" % err_e


@pytest.mark.parametrize("mode,tol", [("f16", 2e-3), ("split", 1e-3), ("bf16", 1e-2)])
def test_outer_conv_operand_modes(cuda_device, mode, tol):
    """The three operand formats of the six outer convs against the fp32 oracle (default-init weights: the
    hard case): fp16 single term (default), split-bf16 hi/lo (three terms), plain bf16."""
    wts = synth.make_weights("default", seed=3, nb=2)
    lr, z = synth.make_inputs(2, 24, 20, seed=3)
    mi = concat_latent(lr, z)
    netG = build_product_G(cuda_device, 2, "all_layers_HR_downscaled", wts)
    netG.generated_image_model.outer_mode = mode
    with torch.no_grad():
        out = netG(mi.to(cuda_device)).cpu()
        ref = GCEMOracle(wts, nb=2).forward(mi)
    assert netG.generated_image_model.engine().outer_mode == mode
    err = (out - ref).abs().max().item()
    assert err <= tol, "mode %s: max|err| %g" % (mode, err)


def test_consistency_and_fp32_oracle_larger_image(cuda_device):
    """1x3x64x64 (BASELINE config 1): parity vs the fp32 oracle and the CEM consistency residual."""
    wts = synth.make_weights("default", seed=0)
    lr, z = synth.make_inputs(1, 64, 64, seed=0)
    mi = concat_latent(lr, z)
    netG = build_product_G(cuda_device, 23, "all_layers_HR_downscaled", wts)
    with torch.no_grad():
        out = netG(mi.to(cuda_device))
        down = netG.DownscaleOP(out)
    res = (down.cpu() - lr).abs()[:, :, 3:-3, 3:-3].max().item()
    assert res <= 1e-4, "consistency residual %g" % res
    with torch.no_grad():
        ref = GCEMOracle(wts).forward(mi)
    err, p = (out.cpu() - ref).abs().max().item(), psnr(out.cpu(), ref)
    assert err <= 1e-2 and p >= 50.0, "max|err| %g, PSNR %.1f dB" % (err, p)


def test_production_net_non_square_input(cuda_device):
    """nb=23, 1x3x72x100 (non-square, not a multiple of the 8x30 tile): PSNR / max error vs the fp32 oracle."""
    wts = synth.make_weights("default", seed=8)
    lr, z = synth.make_inputs(1, 72, 100, seed=8)
    mi = concat_latent(lr, z)
    netG = build_product_G(cuda_device, 23, "all_layers_HR_downscaled", wts)
    with torch.no_grad():
        out = netG(mi.to(cuda_device)).cpu()
        ref = GCEMOracle(wts).forward(mi)
    err, p = (out - ref).abs().max().item(), psnr(out, ref)
    assert out.shape == (1, 3, 288, 400)
    assert err <= 1e-2 and p >= 50.0, "max|err| %g, PSNR %.1f dB" % (err, p)


def test_full_size_batch_config2_properties(cuda_device):
    """BASELINE config 2 at full size (16 x 3x128x128 LR, eval / pre-pad): size-independent properties over the whole
    batch - CEM consistency of every image, image k of the batch == the same image run alone (bit for bit) - and
    PSNR / max error of one image against the fp32 oracle (the oracle takes seconds per image at this size)."""
    wts = synth.make_weights("default", seed=0)
    lr, z = synth.make_inputs(16, 128, 128, seed=21)
    mi = concat_latent(lr, z)
    netG = build_product_G(cuda_device, 23, "all_layers_HR_downscaled", wts)
    with torch.no_grad():
        out = netG(mi.to(cuda_device))
        res = (netG.DownscaleOP(out).cpu() - lr).abs()[:, :, 3:-3, 3:-3]
        alone = netG(mi[9:10].contiguous().to(cuda_device))
    assert out.shape == (16, 3, 512, 512)
    assert res.amax(dim=(1, 2, 3)).max().item() <= 1e-4
    assert torch.equal(alone[0], out[9])
    with torch.no_grad():
        ref = GCEMOracle(wts).forward(mi[4:5])
    err, p = (out[4:5].cpu() - ref).abs().max().item(), psnr(out[4:5].cpu(), ref)
    assert err <= 1e-2 and p >= 50.0, "max|err| %g, PSNR %.1f dB" % (err, p)


def test_batch_shards_are_bit_identical(cuda_device):
    """batch sharding is exact - image i of a batch == the same image run alone."""
    wts = synth.make_weights("kaiming", seed=2, nb=2)
    lr, z = synth.make_inputs(3, 20, 24, seed=2)
    mi = concat_latent(lr, z).to(cuda_device)
    netG = build_product_G(cuda_device, 2, "all_layers_HR_downscaled", wts)
    with torch.no_grad():
        full = netG(mi).clone()
        for i in range(3):
            one = netG(mi[i:i + 1].contiguous())
            assert torch.equal(one[0], full[i])


@pytest.mark.parametrize("B,h,w,chunk,train", [(1, 16, 20, 0, False), (8, 30, 41, 4, False), (6, 24, 70, 2, False),
                                               (3, 52, 33, 1, False),
                                               # unpadded (train-mode) plans with one tile column / one tile pair per
                                               # chunk: the sizes the round-1 fused launch faulted on (divisor-1 magic)
                                               (1, 12, 14, 0, True), (1, 22, 24, 0, True), (1, 33, 35, 0, True),
                                               (2, 8, 8, 1, True), (1, 4, 30, 0, True)])
def test_fused_growth_convs_bit_identical_to_separate_launches(cuda_device, monkeypatch, B, h, w, chunk, train):
    """conv 0..3 of every RDB as ONE persistent launch with tile-level dependencies (esr_rdb_growth_tc) against the
    same convs as four launches: several chunks, several tiles per cluster, odd tile counts; run twice (the flag
    thirds rotate and are cleared by the launches themselves)."""
    wts = synth.make_weights("default", seed=5, nb=2)
    lr, z = synth.make_inputs(B, h, w, seed=5)
    mi = concat_latent(lr, z).to(cuda_device)
    netG = build_product_G(cuda_device, 2, "all_layers_HR_downscaled", wts, train=train)
    G = netG.generated_image_model
    monkeypatch.setenv("ESR_RDB_CHUNK", str(chunk))
    monkeypatch.setenv("ESR_FUSE_RDB", "1")
    with torch.no_grad():
        fused = [netG(mi).clone() for _ in range(3)]
        assert list(G._plans.values())[-1].fuse_rdb
        monkeypatch.setenv("ESR_FUSE_RDB", "0")
        G._plans.clear()
        ref = netG(mi).clone()
        assert not list(G._plans.values())[-1].fuse_rdb
    for f in fused:
        assert torch.equal(f, ref)


def test_tile_sharding_error_vs_halo(cuda_device):
    """parallel.run_tiled: a 1x3x96x96 image as 2 x 2 halo-overlapped tiles through the production net
    against the same image run whole.  The error falls with the halo; at 16 LR px it is far below the parity
    tolerance (the survey's fp32 probe: <= 1e-6; here bf16 activations can flip a rounding, hence 2.5e-3) and the
    stitched image still meets PSNR >= 50 dB / 1e-2 against the fp32 oracle."""
    from esr_b200.parallel import run_tiled
    wts = synth.make_weights("default", seed=0)
    lr, z = synth.make_inputs(1, 96, 96, seed=14)
    mi = concat_latent(lr, z).to(cuda_device)
    netG = build_product_G(cuda_device, 23, "all_layers_HR_downscaled", wts)
    with torch.no_grad():
        whole = netG(mi)
    errs = {}
    for halo in (2, 8, 16):
        tiled = run_tiled(netG, mi, tiles=(2, 2), halo=halo)
        assert tiled.shape == whole.shape
        errs[halo] = (tiled - whole).abs().max().item()
    # measured on B200: 2.3e-2 at halo 2, 1.2e-3 at halo 8 and 16 (the floor is bf16 rounding flips, not the receptive field)
    assert errs[16] <= 2.5e-3 and errs[16] < errs[2] and errs[8] < errs[2], errs
    assert psnr(tiled, whole) >= 60.0
    with torch.no_grad():
        ref = GCEMOracle(wts).forward(mi.cpu())
    err, p = (tiled.cpu() - ref).abs().max().item(), psnr(tiled.cpu(), ref)
    assert err <= 1e-2 and p >= 50.0, "max|err| %g, PSNR %.1f dB" % (err, p)


def test_hilo_trunk_matches_fp32_trunk(cuda_device, monkeypatch):
    """ESR_TRUNK_HILO=1: the residual between the RDBs of an RRDB as a bf16 hi/lo pair (~17 significant bits)."""
    wts = synth.make_weights("default", seed=6, nb=2)
    lr, z = synth.make_inputs(2, 20, 33, seed=6)
    mi = concat_latent(lr, z).to(cuda_device)
    netG = build_product_G(cuda_device, 2, "all_layers_HR_downscaled", wts)
    G = netG.generated_image_model
    with torch.no_grad():
        ref = netG(mi).clone()
        monkeypatch.setenv("ESR_TRUNK_HILO", "1")
        G._plans.clear()
        out = netG(mi).clone()
        assert list(G._plans.values())[-1].trunk_hilo
    assert (out - ref).abs().max().item() < 2e-4


def test_host_pipeline_matches_direct_call(cuda_device):
    """parallel.HostPipeline (chunked, copies overlapped with compute) returns what netG(x) returns."""
    from esr_b200.parallel import HostPipeline
    wts = synth.make_weights("kaiming", seed=4, nb=1)
    lr, z = synth.make_inputs(5, 16, 20, seed=4)
    mi = concat_latent(lr, z).contiguous()
    netG = build_product_G(cuda_device, 1, "all_layers_HR_downscaled", wts)
    with torch.no_grad():
        ref = netG(mi.to(cuda_device)).cpu()
    host_in, host_out = mi.pin_memory(), torch.empty(5, 3, 64, 80).pin_memory()
    pipe = HostPipeline(netG, chunk=2)
    for _ in range(2):                                      # back-to-back calls pipeline as well
        pipe(host_in, host_out)
    pipe.wait()
    assert torch.equal(host_out, ref)


def test_no_cpu_fallback():
    wts = synth.make_weights("kaiming", seed=2, nb=1)
    netG = build_product_G(torch.device("cpu"), 1, "all_layers_HR_downscaled", wts)
    lr, z = synth.make_inputs(1, 8, 8)
    with pytest.raises(Exception):
        netG(concat_latent(lr, z))


@pytest.mark.parametrize("name", ["x2_nb2_eval", "x2_nb1_train"])
def test_x2_generator_matches_reference_golden(golden, cuda_device, name):
    """x2 generator + x2 CEM against the unmodified reference: output (max error <= 1e-2, PSNR >= 50 dB) and the Z
    gradient of its autograd (bf16 dgrad operands: a few per cent)."""
    g = golden("g_cem_x2")
    nb, seed, B, h, w, train = [int(v) for v in g[name + "_cfg"]]
    wts = synth.make_weights(str(g[name + "_kind"]), seed=seed, nb=nb, upscale=2)
    lr, z = synth.make_inputs(B, h, w, sf=2, seed=seed)
    netG = build_product_G(cuda_device, nb, "all_layers_HR_downscaled", wts, train=bool(train), sf=2)
    zp = z.clone().to(cuda_device).requires_grad_(True)
    out = netG(concat_latent(lr.to(cuda_device), zp, sf=2))
    ref = torch.from_numpy(g[name + "_out"])
    assert out.shape == ref.shape
    assert (out.detach().cpu() - ref).abs().max().item() <= 1e-2
    assert psnr(out.detach().cpu(), ref) >= 50.0
    (out * torch.from_numpy(g[name + "_gout"]).to(cuda_device)).sum().backward()
    got, gz = zp.grad.cpu(), torch.from_numpy(g[name + "_gz"])
    assert float((got - gz).norm() / gz.norm()) < 4e-2 and float((got * gz).sum() / (got.norm() * gz.norm())) > 0.999


@pytest.mark.parametrize("name", ["lr_all_nb2_train", "lr_all_nb1_eval", "lr_first_nb1_train"])
def test_lr_domain_latent_matches_reference_golden(golden, cuda_device, name):
    """``latent_input_domain: "LR"`` (architecture.py:137-139,159,165-166): Z of the LR image's size
    assigned to ``.Z`` (SRRaGAN_model.py:260-261), forward on the bare image; all_layers (nearest-upsampled Z at the HR
    convs) and first_layer; train and eval mode (eval: Z at the CEM-padded size).  Against the unmodified reference:
    output max error <= 1e-2 / PSNR >= 50 dB, Z gradient of its autograd within the dgrad tolerance."""
    g = golden("g_cem_lr_domain")
    nb, seed, train = [int(v) for v in g[name + "_cfg"]]
    latent = str(g[name + "_latent"])
    wts = synth.make_weights(str(g[name + "_kind"]), seed=seed, nb=nb, latent_input=latent + "_HR_downscaled")
    netG = build_product_G(cuda_device, nb, latent + "_HR_downscaled", wts, train=bool(train), domain="LR")
    G = netG.generated_image_model
    assert G.latent_input == latent + "_LR" and G.num_latent_channels == 3
    assert hasattr(G, "latent_upsampler") == (latent == "all_layers")
    lr = torch.from_numpy(g[name + "_lr"]).to(cuda_device)
    with pytest.raises(AttributeError):
        netG(lr)                                                    # no Z assigned yet
    zp = torch.from_numpy(g[name + "_z"]).to(cuda_device).requires_grad_(True)
    G.Z = zp
    out = netG(lr)
    ref = torch.from_numpy(g[name + "_out"])
    assert out.shape == ref.shape
    assert (out.detach().cpu() - ref).abs().max().item() <= 1e-2
    assert psnr(out.detach().cpu(), ref) >= 50.0
    (out * torch.from_numpy(g[name + "_gout"]).to(cuda_device)).sum().backward()
    got, gz = zp.grad.cpu(), torch.from_numpy(g[name + "_gz"])
    assert float((got - gz).norm() / gz.norm()) < 5e-2 and float((got * gz).sum() / (got.norm() * gz.norm())) > 0.999
    G.Z = zp.detach()[..., 1:, :]
    with pytest.raises(RuntimeError, match="does not match"):
        netG(lr)


@pytest.mark.parametrize("name", ["rearr_nb2_train", "rearr_nb1_eval"])
def test_hr_rearranged_first_layer_matches_reference_golden(golden, cuda_device, name):
    """``latent_input: "first_layer"`` with ``latent_input_domain: "HR_rearranged"`` (architecture.py:109-110,159): 48
    rearranged latent channels at LR size in ``.Z``, concatenated to the first conv only (the first conv then runs like the
    outer convs: Z as a 64-channel 16-bit tensor + the image as the row-expanded block).  Forward only."""
    g = golden("g_cem_rearranged")
    nb, seed, train = [int(v) for v in g[name + "_cfg"]]
    wts = synth.make_weights(str(g[name + "_kind"]), seed=seed, nb=nb, latent_input="first_layer_HR_rearranged", num_latent_channels=48)
    netG = build_product_G(cuda_device, nb, "first_layer_HR_downscaled", wts, train=bool(train), domain="HR_rearranged")
    G = netG.generated_image_model
    assert G.latent_input == "first_layer_HR_rearranged" and G.num_latent_channels == 48
    lr = torch.from_numpy(g[name + "_lr"]).to(cuda_device)
    G.Z = torch.from_numpy(g[name + "_z"]).to(cuda_device)
    with torch.no_grad():
        out = netG(lr)
    ref = torch.from_numpy(g[name + "_out"])
    assert out.shape == ref.shape
    assert (out.cpu() - ref).abs().max().item() <= 1e-2
    assert psnr(out.cpu(), ref) >= 50.0
    G.Z = G.Z.clone().requires_grad_(True)
    with pytest.raises(NotImplementedError, match="forward only"):
        netG(lr)


def test_pretrained_checkpoint_without_latent_is_reproduced(cuda_device):
    """base_model.py:126-136: a generator without Z loaded into the Z-conditioned one (zero weights for the new input
    channels) must output what the original does, whatever Z is - until the Z weights are trained."""
    from esr_b200 import checkpoint
    from tests.test_checkpoint import esrgan_style_checkpoint, make_opt
    ckpt, plain_w = esrgan_style_checkpoint(nb=2, seed=12)
    lat = networks.define_G(make_opt(nb=2), CEM=pcem.CEMnet(pcem.Get_CEM_Config(4)), num_latent_channels=3)
    checkpoint.load_network(ckpt, lat, latent_input="all_layers_HR_downscaled", num_latent_channels=3)
    lat = lat.to(cuda_device).eval()
    plain = build_product_G(cuda_device, 2, None, plain_w)
    lr, z = synth.make_inputs(1, 18, 14, seed=12)
    with torch.no_grad():
        a = lat(concat_latent(lr, z).to(cuda_device))
        b = lat(concat_latent(lr, -z).to(cuda_device))
        ref = plain(lr.to(cuda_device))
    assert torch.equal(a, b)
    # the two engines accumulate in a different order (extra latent K blocks contributing exact zeros, other slot
    # layout of the first conv): fp32 summation-order noise, 4e-5 on outputs of O(1)
    assert (a - ref).abs().max().item() <= 2e-4
