"""The critic of the GAN training step (BASELINE config 5).

``Discriminator_VGG_128_`` mirrors codes/models/modules/architecture.py:222-284 — the class the reference's training
configuration means (`network_D.n_layers`, `num_2_strides`): alternating 3x3 stride-1 / 4x4 stride-2 convolutions with
BatchNorm + LeakyReLU(0.2) cut after `nb` layers, then a patch head (8x8 valid conv to min(100, C) channels, BatchNorm,
two LeakyReLUs, 1x1 conv to one channel, BatchNorm, LeakyReLU).  Module tree and state-dict keys are the reference's
(`features.{i}.weight`, `classifier.0.{0,1}.*`, `classifier.2.{0,1}.*`), so its checkpoints load.

``define_D`` mirrors codes/models/networks.py:105-127 with the one defect of the shipped code routed around (SURVEY.md
§8a "Reference bugs"): it names ``Discriminator_VGG_128``, whose constructor does not take the ``nb=`` it is given
(networks.py:119 vs architecture.py:182), so the reference cannot construct its own critic; the class that takes
`nb` / `num_2_strides` is ``Discriminator_VGG_128_``.

Arithmetic: the critic is 1.6 % of the training step's FLOPs and needs a double backward (WGAN-GP, loss.py:244-263), so
it runs as torch modules under torch autograd (cuDNN: LIBRARY code, declared as such in bench.py's `train_gan` line); the
generator's forward, data gradient and weight gradients under it are this package's kernels (rrdbnet._TrainFn)."""
import math

import torch.nn as nn

from .networks import _SingleDeviceParallel, init_weights


def _layer(cin, cout, k, stride, norm, pad=None):
    mods = [nn.Conv2d(cin, cout, k, stride, (k - 1) // 2 if pad is None else pad, bias=True)]
    if norm:
        mods.append(nn.BatchNorm2d(cout, affine=True))
    mods.append(nn.LeakyReLU(0.2, True))
    return mods


class Discriminator_VGG_128_(nn.Module):
    def __init__(self, in_nc, base_nf, norm_type='batch', act_type='leakyrelu', mode='CNA', input_patch_size=128,
                 num_2_strides=5, nb=10):
        super().__init__()
        if act_type != 'leakyrelu' or mode != 'CNA' or norm_type not in ('batch', None):
            raise NotImplementedError("Discriminator_VGG_128_: act_type 'leakyrelu', mode 'CNA', norm_type 'batch' / None")
        if num_2_strides > 5:
            raise AssertionError('Can be modified by adding more stridable layers, if needed.')
        self.num_2_strides = num_2_strides
        self.last_FC_layers = False                        # architecture.py:268-269: always the patch head
        widths = [base_nf, base_nf, 2 * base_nf, 2 * base_nf, 4 * base_nf, 4 * base_nf, 8 * base_nf, 8 * base_nf, 8 * base_nf,
                  8 * base_nf]
        layers, cin, size, strides_left = [], in_nc, input_patch_size, num_2_strides
        for i, cout in enumerate(widths[:nb]):
            if i % 2 == 0:                                 # 3x3, stride 1; the very first one without a norm
                layers += _layer(cin, cout, 3, 1, norm_type is not None and i > 0)
            else:                                          # 4x4, stride 2 while strides remain
                s = 2 if strides_left > 0 else 1
                layers += _layer(cin, cout, 4, s, norm_type is not None)
                size = math.ceil((size - 1) / s)
                strides_left -= 1
            cin = cout
        self.features = nn.Sequential(*layers)
        self.feature_size = size
        norms = [m for m in self.features if isinstance(m, nn.BatchNorm2d)]
        if not norms or not isinstance(self.features[-2], nn.BatchNorm2d):
            raise AttributeError("the patch head takes its width from the last feature layer's norm (architecture.py:275)")
        c_feat, c_mid = self.features[-2].num_features, min(100, self.features[-2].num_features)
        self.classifier = nn.Sequential(nn.Sequential(*_layer(c_feat, c_mid, 8, 1, True, pad=0)), nn.LeakyReLU(0.2, False),
                                        nn.Sequential(*_layer(c_mid, 1, 1, 1, True)))

    def forward(self, x):
        return self.classifier(self.features(x))


def define_D(opt, CEM=None):
    gpu_ids = opt['gpu_ids']
    opt_net = opt['network_D']
    which_model = opt_net['which_model_D']
    input_patch_size = opt['datasets']['train']['patch_size']
    if (opt_net['pre_clipping'] or opt_net['decomposed_input']) and which_model != 'PatchGAN':
        raise AssertionError('Unsupported yet')
    if CEM is not None:
        input_patch_size -= 2 * CEM.invalidity_margins_HR
    if which_model != 'discriminator_vgg_128':
        raise NotImplementedError('Discriminator model [{:s}] not recognized'.format(which_model))
    kwargs = {'num_2_strides': opt_net['num_2_strides']} if 'num_2_strides' in opt_net and opt_net['num_2_strides'] is not None else {}
    netD = Discriminator_VGG_128_(in_nc=opt_net['in_nc'], base_nf=opt_net['nf'], nb=opt_net['n_layers'], norm_type=opt_net['norm_type'],
                                  mode=opt_net['mode'], act_type=opt_net['act_type'], input_patch_size=input_patch_size, **kwargs)
    init_weights(netD, init_type='kaiming', scale=1)
    if gpu_ids:
        netD = _DiscriminatorParallel(netD)
    return netD


class _DiscriminatorParallel(_SingleDeviceParallel):
    """`.module` without DataParallel's scatter (one process per GPU).  The class name matters: the reference's
    ``get_network_description`` (base_model.py:48-58) only unwraps ``nn.DataParallel`` and returns the receptive field
    for classes whose name contains 'Discriminator'."""
