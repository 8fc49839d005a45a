"""CPU-only tests of the host side: filter derivation, module tree / state_dict contract, engine layout
arithmetic, C-ABI library exports.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from esr_b200 import _capi as capi, cem as pcem, networks, synth
from esr_b200.engine import GEngine, expand_slots
from oracle import cem_filters

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_opt(nb=23, latent="all_layers", is_train=False):
    return {"gpu_ids": None, "is_train": is_train, "datasets": {"train": {"patch_size": 256}},
            "network_G": dict(which_model_G="RRDB_net", CEM_arch=1, latent_input=latent, latent_input_domain="HR_downscaled",
                              latent_channels=3, norm_type=None, mode="CNA", nf=64, nb=nb, in_nc=3, out_nc=3, gc=32, scale=4)}


@pytest.mark.parametrize("sf", [2, 3, 4])
def test_product_filters_match_oracle_and_golden(golden, sf):
    c = pcem.CEMnet(pcem.Get_CEM_Config(sf))
    g = golden("cem_filters")
    np.testing.assert_allclose(c.ds_kernel, g["ds_kernel_%d" % sf], atol=1e-7)
    np.testing.assert_allclose(c.inv_hTh, g["inv_hTh_%d" % sf], atol=2e-6)
    assert [c.invalidity_margins_LR, int(c.invalidity_margins_HR)] == list(g["margins_%d" % sf][:2])
    o = cem_filters.derive(sf)
    np.testing.assert_allclose(np.outer(c._ds_1d, c._ds_1d), o["ds_kernel"], atol=1e-7)
    np.testing.assert_allclose(np.outer(c._inv_1d, c._inv_1d), o["inv_hTh"], atol=2e-6)


def test_state_dict_contract():
    """Key names, order and shapes of the reference module tree, 17 060 948 + 3 921 parameters."""
    netG = networks.define_G(make_opt(), CEM=pcem.CEMnet(pcem.Get_CEM_Config(4)), num_latent_channels=3)
    sd = netG.state_dict()
    w = synth.make_weights("default")
    assert [k for k in sd if "Filter" not in k] == ["generated_image_model." + k for k in w]
    for k, v in w.items():
        assert sd["generated_image_model." + k].shape == v.shape
    assert sd["Conv_LR_with_Inv_hTh_OP.Filter_OP.weight"].shape == (3, 1, 27, 27)
    assert sd["Upscale_OP.Filter_OP.weight"].shape == (3, 1, 17, 17) and sd["DownscaleOP.Filter_OP.weight"].shape == (3, 1, 17, 17)
    assert sum(p.numel() for p in netG.parameters() if p.requires_grad) == 17060948
    assert sum(p.numel() for p in netG.parameters() if not p.requires_grad) == 3921
    assert sd["generated_image_model.model.1.sub.0.RDB1.convs.0.0.weight"].shape == (32, 67, 3, 3)
    assert netG.pre_pad is False and netG.eval().pre_pad is True and netG.train().pre_pad is False


def test_define_G_modes_and_init():
    torch.manual_seed(0)
    cem = pcem.CEMnet(pcem.Get_CEM_Config(4))
    g_train = networks.define_G(make_opt(nb=1, is_train=True), CEM=cem, num_latent_channels=3)
    assert all(float(p.abs().max()) == 0 for n, p in g_train.named_parameters() if n.endswith("bias"))
    k = g_train.Upscale_OP.Filter_OP.weight.clone()          # kaiming init must skip the CEM filters
    np.testing.assert_allclose(k[0, 0].numpy(), cem.ds_kernel * 16, atol=1e-7)
    opt = make_opt(nb=1, latent="None")
    g_plain = networks.define_G(opt, CEM=cem, num_latent_channels=0)
    assert opt["network_G"]["latent_input"] is None
    assert g_plain.generated_image_model.model[0].weight.shape == (64, 3, 3, 3)
    g_first = networks.define_G(make_opt(nb=1, latent="first_layer"), CEM=cem, num_latent_channels=3)
    assert g_first.generated_image_model.model[0].weight.shape == (64, 6, 3, 3)
    assert g_first.generated_image_model.model[1].sub[0].RDB1.convs[0][0].weight.shape == (32, 64, 3, 3)
    wrapped = networks.define_G({**make_opt(nb=1), "gpu_ids": [0]}, CEM=cem, num_latent_channels=3) if torch.cuda.is_available() else None
    assert wrapped is None or hasattr(wrapped, "module")


def test_cpu_forward_fails_loudly():
    netG = networks.define_G(make_opt(nb=1), CEM=pcem.CEMnet(pcem.Get_CEM_Config(4)), num_latent_channels=3).eval()
    for p in netG.parameters():
        p.requires_grad_(False)
    with pytest.raises(Exception, match="CUDA|CPU"):
        netG(torch.zeros(1, 51, 8, 8))
    with pytest.raises(Exception, match="CUDA|CPU"):
        netG.DownscaleOP(torch.zeros(1, 3, 16, 16))


def test_engine_layout_arithmetic():
    e = GEngine(nb=23, nz_in=3, all_layers=True)
    assert len(e.convs) == 351
    for pc in e.convs.values():
        pc.cin = None
    shapes = synth.rrdbnet_conv_shapes()
    for name, (co, ci) in shapes.items():
        e.convs[name].cin = ci
        assert e.convs[name].cout == co
    assert e.flops_per_lr_pixel() == 2 * 18316944               # 36.634 MFLOP per LR pixel
    c0 = e.convs["model.1.sub.0.RDB1.convs.0.0"]
    assert (c0.nkb, c0.cout_tiles, c0.w_tile_bytes) == (3, 1, 2 * 18432 + 3072)      # lean latent block: 16-channel slabs
    assert [k.half for k in list(c0.kblocks)[:3]] == [0, 0, 1]
    c4 = e.convs["model.1.sub.0.RDB1.convs.4.0"]
    assert (c4.nkb, c4.cout_tile, c4.cout_tiles, c4.pair) == (7, 64, 1, 1) and c4.w_tile_bytes // 2 <= 124 * 1024
    # outer convs, default mode: fp16 operands, one term (2 main K blocks + the fp16 latent slice)
    lr_conv = e.convs["model.1.sub.23"]
    assert e.outer_mode == "f16" and "model.1.sub.23" in e.f16_convs and "model.1.sub.0.RDB1.convs.0.0" not in e.f16_convs
    assert lr_conv.nkb == 3 and [(k.chan, k.slice_mask) for k in list(lr_conv.kblocks)[:3]] == [(0, 3), (32, 3), (0, 2)]
    assert all(sl.term == 2 for sl in list(lr_conv.slots)[:64])
    # split-bf16 mode: hi/lo pairs, three MMA terms
    es = GEngine(nb=23, nz_in=3, all_layers=True, outer_mode="split")
    lr_conv = es.convs["model.1.sub.23"]
    assert lr_conv.nkb == 8 and [k.chan for k in list(lr_conv.kblocks)[:6]] == [0, 32, 64, 96, 0, 32]
    xs, ws = expand_slots(6, precise=True)
    assert len(xs) == 64 and xs[0] == (0, -1, 0) and xs[18] == (0, -1, 1) and ws[36] == (0, 0, 1)
    xs, ws = expand_slots(6, precise="f16")
    assert len(xs) == 32 and xs[0] == (0, -1, 2) and ws[17] == (5, 2, 2) and xs[18] == (-1, 0, 0)
    xs, ws = expand_slots(3, precise=False, second="f16")
    assert len(xs) == 32 and xs[8] == (2, 1, 0) and xs[16] == (0, -1, 2) and ws[16] == (0, 0, 2) and ws[9] == (-1, -1, 0)
    # LR_conv must not be mistaken for an upconv when nb == 1 ("model.1.sub.1")
    e1 = GEngine(nb=1, nz_in=3, all_layers=True)
    assert e1.convs["model.1.sub.1"].nkb == 3 and "model.1.sub.1" not in e1.upconv_names


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "esr_b200.h")).read()
    declared = set(re.findall(r"\b(esr_[a-z0-9_]+)\s*\(", header))
    declared -= {"esr_conv_desc", "esr_kblock"}
    lib = capi.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), "libesr_b200.so does not export %s" % name
    assert set(capi.SIGNATURES) <= declared
    assert lib.esr_abi_version() == 2
    assert ctypes.sizeof(capi.ConvDesc) % 8 == 0
    # host-only entry point: weight image layout needs no GPU
    kb = (capi.KBlock * capi.MAX_KBLOCKS)()
    kb[0].dy_mask, kb[0].slice_mask = 0b111, 0b11
    kb[1].dy_mask, kb[1].slice_mask = 0b010, 0b01
    wtb = ctypes.c_uint32(0)
    total = lib.esr_pack_layout(32, 2, 0, 2, kb, ctypes.byref(wtb))
    assert (wtb.value, total, kb[1].w_off, kb[0].n_dy, kb[1].n_dy) == (18432 + 6144, 2 * 24576, 18432, 3, 1)
    # pair layout: same bytes per cout tile, split into two half images; w_off is relative to a half
    total = lib.esr_pack_layout(32, 2, 1, 2, kb, ctypes.byref(wtb))
    assert (wtb.value, total, kb[1].w_off) == (18432 + 6144, 2 * 24576, 9216)
    total = lib.esr_pack_layout(64, 1, 1, 2, kb, ctypes.byref(wtb))
    assert (wtb.value, total, kb[1].w_off) == (2 * 24576, 2 * 24576, 18432)
    assert lib.esr_pack_layout(64, 1, 0, 2, kb, ctypes.byref(wtb)) < 0          # 64-channel tiles need pair mode
    assert lib.esr_pack_layout(7, 1, 0, 1, kb, ctypes.byref(wtb)) < 0 and b"esr_pack_layout" in lib.esr_last_error()


def test_new_entry_points_validate_arguments_without_a_gpu():
    """Argument validation of the round-2 entry points happens before any CUDA call: bad arguments return ESR_ERR_INVALID
    with a message, struct mirrors match the header's layout (sizes the kernels index by)."""
    lib = capi.lib()
    assert ctypes.sizeof(capi.WgradTcItem) == 64 and ctypes.sizeof(capi.WgradItem) % 8 == 0
    assert lib.esr_wgrad_tc_map_bytes() == 128
    assert lib.esr_wgrad_tc(None, 0, None, None) == -1 and b"esr_wgrad_tc" in lib.esr_last_error()
    assert lib.esr_wgrad16r(None, 1, None) == -1
    buf = (ctypes.c_uint8 * 128)()
    assert lib.esr_wgrad_tc_make_map(buf, None, 64, 1, 8, 16, 0) == -1            # null base
    assert lib.esr_wgrad_tc_make_map(buf, ctypes.c_void_p(4096), 60, 1, 8, 16, 0) == -1    # channels % 8
    assert lib.esr_wgrad_tc_make_map(buf, ctypes.c_void_p(4096), 64, 1, 8, 16, 2) == -1    # kind
    f = capi.cem_filters_struct(4, 1, [0.0] * 17, [0.0] * 27)
    assert lib.esr_cem_project_fused(f, None, None, 1, 3, 64, 1024, 0, None, None, None) == -1
    assert b"esr_cem_project_fused" in lib.esr_last_error()
    assert lib.esr_cem_project_fused(f, ctypes.c_void_p(4096), ctypes.c_void_p(4096), 1, 3, 64, 1024, 40, ctypes.c_void_p(4096),
                                     ctypes.c_void_p(4096), None) == -1       # crop too large for H = 64
