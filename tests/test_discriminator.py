"""The critic of the training step: module tree, state-dict keys and outputs of
esr_b200.discriminator.Discriminator_VGG_128_ against the reference's class (architecture.py:222-284), the parameter
count SURVEY.md §8e quotes, and define_D on the training options (networks.py:105-127)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from esr_b200 import discriminator as pd  # noqa: E402


def test_parameter_count_and_patch_head():
    d = pd.Discriminator_VGG_128_(3, 64, nb=6)
    assert sum(p.numel() for p in d.parameters()) == 3387987        # 13.6 MB of gradients per all-reduce
    assert d(torch.rand(2, 3, 128, 128)).shape == (2, 1, 9, 9)
    assert d(torch.rand(1, 3, 176, 176)).shape == (1, 1, 15, 15)    # patch 256 minus the CEM margins (train JSON)
    with pytest.raises(AssertionError):
        pd.Discriminator_VGG_128_(3, 64, num_2_strides=6)


@pytest.mark.parametrize("nb,strides,patch", [(6, 5, 128), (10, 5, 256), (6, 2, 96), (4, 5, 64)])
def test_equals_reference_class(nb, strides, patch):
    from oracle import ref_shims
    if not ref_shims.available():
        pytest.skip("reference tree not available")
    arch = ref_shims.load_reference()[2]
    torch.manual_seed(0)
    want = arch.Discriminator_VGG_128_(3, 16, nb=nb, num_2_strides=strides, input_patch_size=patch)
    have = pd.Discriminator_VGG_128_(3, 16, nb=nb, num_2_strides=strides, input_patch_size=patch)
    assert list(want.state_dict().keys()) == list(have.state_dict().keys())
    have.load_state_dict(want.state_dict())
    x = torch.rand(2, 3, patch, patch)
    assert torch.equal(want(x), have(x))                            # train mode: batch statistics
    want.eval(), have.eval()
    assert torch.equal(want(x), have(x))


def test_define_d_on_training_options():
    class CEM:
        invalidity_margins_HR = 40
    opt = {"gpu_ids": None, "datasets": {"train": {"patch_size": 256}},
           "network_D": {"which_model_D": "discriminator_vgg_128", "in_nc": 3, "nf": 8, "n_layers": 6, "norm_type": "batch", "mode": "CNA",
                         "act_type": "leakyrelu", "pre_clipping": 0, "decomposed_input": 0}}
    d = pd.define_D(opt, CEM=CEM)
    assert isinstance(d, pd.Discriminator_VGG_128_) and d.feature_size == 22
    assert all(float(m.bias.abs().max()) == 0 for m in d.modules() if isinstance(m, torch.nn.Conv2d))
    opt["network_D"]["which_model_D"] = "PatchGAN"
    with pytest.raises(NotImplementedError):
        pd.define_D(opt)
