"""Network factory.  Drop-in for ``models.networks.define_G`` / ``init_weights``
(codes/models/networks.py:28-102) restricted to the generator path."""
import functools

import torch
import torch.nn as nn
from torch.nn import init

from .rrdbnet import RRDBNet


def weights_init_kaiming(m, scale=1):
    if getattr(m, 'filter_layer', False):           # CEM filters stay fixed (networks.py:29-30)
        return
    name = m.__class__.__name__
    if name.find('Conv') != -1 or name.find('Linear') != -1:
        init.kaiming_normal_(m.weight.data, a=0, mode='fan_in')
        m.weight.data *= scale
        if m.bias is not None:
            m.bias.data.zero_()
    elif name.find('BatchNorm2d') != -1:
        init.constant_(m.weight.data, 1.0)
        init.constant_(m.bias.data, 0.0)


def init_weights(net, init_type='kaiming', scale=1, std=0.02):
    print('initialization method [{:s}]'.format(init_type))
    if init_type != 'kaiming':
        raise NotImplementedError('initialization method [{:s}] not implemented'.format(init_type))
    net.apply(functools.partial(weights_init_kaiming, scale=scale))


def define_G(opt, CEM=None, num_latent_channels=None):
    gpu_ids = opt['gpu_ids']
    opt_net = opt['network_G']
    which_model = opt_net['which_model_G']
    opt_net['latent_input'] = opt_net['latent_input'] if opt_net['latent_input'] != "None" else None
    if which_model != 'RRDB_net':
        raise NotImplementedError('Generator model [{:s}] not recognized'.format(which_model))
    latent = (opt_net['latent_input'] + '_' + opt_net['latent_input_domain']) if opt_net['latent_input'] is not None else None
    netG = RRDBNet(in_nc=opt_net['in_nc'], out_nc=opt_net['out_nc'], nf=opt_net['nf'], nb=opt_net['nb'],
                   gc=opt_net['gc'], upscale=opt_net['scale'], norm_type=opt_net['norm_type'], act_type='leakyrelu',
                   mode=opt_net['mode'], upsample_mode='upconv', latent_input=latent,
                   num_latent_channels=num_latent_channels)
    if opt_net['CEM_arch']:
        netG = CEM.WrapArchitecture_PyTorch(netG, opt['datasets']['train']['patch_size'] if opt['is_train'] else None)
    if opt['is_train']:
        init_weights(netG, init_type='kaiming', scale=0.1)
    if gpu_ids:
        assert torch.cuda.is_available()
        # One process per GPU (bench.py / torch.distributed) replaces nn.DataParallel's thread-per-GPU
        # scatter (networks.py:99-101); `.module` is kept for callers that reach through it.
        netG = _SingleDeviceParallel(netG)
    return netG


class _SingleDeviceParallel(nn.Module):
    """Keeps the ``netG.module`` attribute path of nn.DataParallel without its scatter/gather."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **kw):
        return self.module(*a, **kw)
