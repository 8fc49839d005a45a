"""Checkpoint import (codes/models/base_model.py:100-144): positional key matching, zero weights for the latent input
channels a pre-trained network did not have, CEM filters never loaded.  Host logic, CPU only; where the reference tree is
present (build container) the result is also compared with the reference's own ``process_loaded_state_dict``."""
import collections
import io
import types

import numpy as np
import pytest
import torch

from esr_b200 import cem as pcem, checkpoint, networks, synth
from oracle import ref_shims


def make_opt(nb=1, latent="all_layers", sf=4):
    return {"gpu_ids": None, "is_train": False, "scale": sf, "datasets": {"train": {"patch_size": 256}},
            "network_G": dict(which_model_G="RRDB_net", CEM_arch=1, latent_input=latent, latent_input_domain="HR_downscaled",
                              latent_channels=3, norm_type=None, mode="CNA", nf=64, nb=nb, in_nc=3, out_nc=3, gc=32, scale=sf)}


def esrgan_style_checkpoint(nb=1, seed=3):
    """A pre-trained generator without Z under the public ESRGAN key naming (RDB1.conv1.0.weight ...)."""
    w = synth.make_weights("default", seed=seed, nb=nb, latent_input=None)
    out = collections.OrderedDict()
    for k, v in w.items():
        out[k.replace(".convs.0.0.", ".conv1.0.").replace(".convs.1.0.", ".conv2.0.").replace(".convs.2.0.", ".conv3.0.")
             .replace(".convs.3.0.", ".conv4.0.").replace(".convs.4.0.", ".conv5.0.")] = v
    return out, w


@pytest.fixture()
def latent_G():
    return networks.define_G(make_opt(), CEM=pcem.CEMnet(pcem.Get_CEM_Config(4)), num_latent_channels=3)


def test_pretrained_without_latent_loads_into_latent_generator(latent_G):
    ckpt, plain = esrgan_style_checkpoint()
    filters_before = {k: v.clone() for k, v in latent_G.state_dict().items() if "Filter" in k}
    amp = {}
    checkpoint.load_network({"model_state_dict": ckpt, "optimizer_state_dict": {}}, latent_G, latent_input="all_layers_HR_downscaled",
                            num_latent_channels=3, scale=4, grad_amplification=amp)
    sd = latent_G.state_dict()
    n_ext = 0
    for (k_old, v), k_new in zip(plain.items(), [k for k in sd if "Filter" not in k]):
        assert k_new == "generated_image_model." + k_old
        t = sd[k_new]
        if t.dim() == 4 and t.shape[1] == v.shape[1] + 3:      # every conv of the all_layers net takes Z first on dim 1
            assert torch.equal(t[:, 3:], v) and float(t[:, :3].abs().max()) == 0.0
            n_ext += 1
        else:
            assert torch.equal(t, v)
    assert n_ext == 1 + 15 + 1 + 2                              # first conv, 15 trunk convs, LR_conv, HR convs (not the upconvs)
    assert len(amp) == n_ext and all(v == [0, 1, 2] for v in amp.values())
    for k, v in filters_before.items():
        assert torch.equal(sd[k], v)


def test_round_trip_and_refusals(latent_G, tmp_path):
    path = checkpoint.save_network(str(tmp_path / "1_G.pth"), latent_G)
    other = networks.define_G(make_opt(), CEM=pcem.CEMnet(pcem.Get_CEM_Config(4)), num_latent_channels=3)
    checkpoint.load_network(path, other, latent_input="all_layers_HR_downscaled", num_latent_channels=3)   # CEM filters are skipped: not strict
    for (k, a), (_, b) in zip(latent_G.state_dict().items(), other.state_dict().items()):
        assert torch.equal(a, b), k
    ckpt, _ = esrgan_style_checkpoint()
    short = collections.OrderedDict(list(ckpt.items())[:-2])
    with pytest.raises(ValueError):
        checkpoint.load_network(short, other, latent_input="all_layers_HR_downscaled", num_latent_channels=3)
    bad = collections.OrderedDict(ckpt)
    first = next(iter(bad))
    bad[first] = torch.zeros(32, 3, 3, 3)                      # wrong Cout under a renamed key
    bad = collections.OrderedDict(("x." + k if k == first else k, v) for k, v in bad.items())
    with pytest.raises(ValueError):
        checkpoint.load_network(bad, other, latent_input="all_layers_HR_downscaled", num_latent_channels=3)


@pytest.mark.skipif(not ref_shims.available(), reason="reference tree not present (GPU box)")
def test_matches_reference_loader(latent_G, monkeypatch):
    ref_shims.install()
    import models.base_model as ref_bm
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    ckpt, _ = esrgan_style_checkpoint()
    current = latent_G.state_dict()
    loaded = pcem.Adjust_State_Dict_Keys(ckpt, current)
    op_names = [n for n, _ in latent_G.named_modules() if "Filter_OP" in n]
    me = types.SimpleNamespace(latent_input="all_layers_HR_downscaled", num_latent_channels=3, opt={"scale": 4},
                               CEM_net=types.SimpleNamespace(OP_names=op_names), CEM_arch=True,
                               channels_idx_4_grad_amplification={})
    ref = ref_bm.BaseModel.process_loaded_state_dict(me, loaded_state_dict=loaded, current_state_dict=current)
    amp = {}
    mine = checkpoint.process_loaded_state_dict(loaded, current, latent_input="all_layers_HR_downscaled", num_latent_channels=3,
                                                scale=4, cem_op_names=op_names, grad_amplification=amp)
    assert list(ref.keys()) == list(mine.keys())
    for k in ref:
        np.testing.assert_allclose(mine[k].numpy(), ref[k].numpy(), atol=0, rtol=0)
    assert amp == me.channels_idx_4_grad_amplification
