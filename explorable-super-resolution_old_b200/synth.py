"""Deterministic synthetic weights and inputs (there is no network for checkpoints).

The generator is independent of torch's RNG stream so that the build container,
the GPU box, the oracle and the CUDA path all see bit-identical fp32 weights
without shipping a 68 MB state_dict: every tensor is drawn from a numpy PCG64
stream seeded by crc32(key) ^ seed.

Two weight sets mirror what ``define_G`` yields (codes/models/networks.py:83-102):
``default``  - torch's Conv2d default (kaiming_uniform(a=sqrt 5): U(+-1/sqrt(fan_in))
               for weight and bias), the test/GUI path without a checkpoint;
``kaiming``  - kaiming_normal(fan_in) * 0.1, zero bias (is_train path, :97-98).
"""
import math
import zlib
from collections import OrderedDict

import numpy as np
import torch


def rrdbnet_conv_shapes(nf=64, nb=23, gc=32, in_nc=3, out_nc=3, upscale=4,
                        latent_input="all_layers_HR_downscaled", num_latent_channels=3):
    """Ordered {state_dict key prefix: (Cout, Cin)} of every 3x3 conv of RRDBNet.

    Key names and order follow the module tree built at
    codes/models/modules/architecture.py:102-150 (latent channels first on dim 1).
    """
    nz_in = num_latent_channels if latent_input is not None else 0
    nz = nz_in if (latent_input is not None and "all_layers" in latent_input) else 0
    n_up = 1 if upscale == 3 else int(math.log(upscale, 2))
    shapes = OrderedDict()
    shapes["model.0"] = (nf, in_nc + nz_in)
    for r in range(nb):
        for d in (1, 2, 3):
            for i in range(5):
                shapes["model.1.sub.%d.RDB%d.convs.%d.0" % (r, d, i)] = (gc if i < 4 else nf, nf + i * gc + nz)
    shapes["model.1.sub.%d" % nb] = (nf, nf + nz)
    for u in range(n_up):
        shapes["model.%d.1" % (2 + u)] = (nf, nf)
    shapes["model.%d" % (2 + n_up)] = (nf, nf + nz)
    shapes["model.%d" % (4 + n_up)] = (out_nc, nf + nz)
    return shapes


def _rng(key, seed):
    return np.random.Generator(np.random.PCG64((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0xFFFFFFFF))


def make_weights(kind="default", seed=0, prefix="", **cfg):
    """fp32 CPU tensors keyed ``<prefix><conv>.weight|bias``."""
    out = OrderedDict()
    for name, (co, ci) in rrdbnet_conv_shapes(**cfg).items():
        fan_in = ci * 9
        if kind == "default":
            b = 1.0 / math.sqrt(fan_in)
            w = _rng(name + ".weight", seed).uniform(-b, b, size=(co, ci, 3, 3))
            bias = _rng(name + ".bias", seed).uniform(-b, b, size=(co,))
        elif kind == "kaiming":
            w = _rng(name + ".weight", seed).standard_normal(size=(co, ci, 3, 3)) * (math.sqrt(2.0 / fan_in) * 0.1)
            bias = np.zeros((co,))
        else:
            raise ValueError(kind)
        out[prefix + name + ".weight"] = torch.from_numpy(w.astype(np.float32))
        out[prefix + name + ".bias"] = torch.from_numpy(bias.astype(np.float32))
    return out


def make_inputs(batch, h, w, sf=4, num_latent_channels=3, seed=0):
    """LR in [0,1], Z in [-1,1] (SRRaGAN_model.py:279,285), fp32 CPU tensors."""
    lr = _rng("LR", seed).random(size=(batch, 3, h, w)).astype(np.float32)
    z = (2.0 * _rng("Z", seed).random(size=(batch, num_latent_channels, sf * h, sf * w)) - 1.0).astype(np.float32)
    return torch.from_numpy(lr), torch.from_numpy(z)
