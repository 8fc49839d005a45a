"""Checkpoint import / export for the generator.  Drop-in for ``BaseModel.load_network`` / ``save_network`` /
``process_loaded_state_dict`` (codes/models/base_model.py:85-144): the data format on the input side of the hot path.

What the reference's loader does, and this one does the same way:
* a checkpoint is either a bare state_dict or ``{'model_state_dict': ..., 'optimizer_state_dict': ...}`` (:103-107);
* a plain-ESRGAN state_dict gets the ``generated_image_model.`` prefix and the current CEM filter entries
  (``CEMnet.Adjust_State_Dict_Keys``, :108-109);
* tensors are matched to the current module BY POSITION, not by name (:118-125) - that is what lets the public
  ``RRDB_ESRGAN_x4.pth`` (``RDB1.conv1.0.weight`` naming) load into the ModuleList naming (``RDB1.convs.0.0.weight``);
  a renamed tensor must agree in every dimension except dim 1;
* a conv that gained latent input channels (the latent channels come FIRST on dim 1, by ``num_latent_channels`` or by
  ``num_latent_channels * scale**2``) keeps the loaded weights for the old channels and gets zeros for the new ones
  (:126-136 with LATENT_WEIGHTS_RELATIVE_STD = 0), so a pre-trained network without Z is reproduced exactly until Z's
  weights are trained; the positions of those convs are reported in ``channels_idx_4_grad_amplification``;
* CEM filter entries of the checkpoint are never loaded (:137-138): the filters follow from the CEM configuration.
Host-side only (state_dict surgery); nothing here touches the GPU.
"""
import collections

import os

import torch
import torch.nn as nn

from .cem import Adjust_State_Dict_Keys


def process_loaded_state_dict(loaded_state_dict, current_state_dict, latent_input=None, num_latent_channels=0, scale=4,
                              cem_op_names=(), grad_amplification=None):
    """base_model.py:113-144.  Returns the state_dict to hand to ``load_state_dict(strict=False)``.
    grad_amplification: optional dict that receives {position: [new input channel indices]}."""
    out = collections.OrderedDict()
    current_keys = list(current_state_dict.keys())
    if len(current_keys) != len(loaded_state_dict):
        raise ValueError("Loaded model and current one should have the same number of parameters (%d vs %d)"
                         % (len(loaded_state_dict), len(current_keys)))
    renamed = 0
    widths = [num_latent_channels, num_latent_channels * scale ** 2] if (latent_input is not None and num_latent_channels) else []
    for i, (key, value) in enumerate(loaded_state_dict.items()):
        cur_key = current_keys[i]
        cur = current_state_dict[cur_key]
        if key != cur_key:
            if tuple(value.shape[:1]) + tuple(value.shape[2:]) != tuple(cur.shape[:1]) + tuple(cur.shape[2:]):
                raise ValueError("Unmatching parameter sizes after changing parameter key name (%s -> %s: %s vs %s)"
                                 % (key, cur_key, tuple(value.shape), tuple(cur.shape)))
            renamed += 1
        if widths and 'weight' in key and value.dim() > 1 and (cur.shape[1] - value.shape[1]) in widths:
            extra = cur.shape[1] - value.shape[1]
            pad = torch.zeros((cur.shape[0], extra) + tuple(cur.shape[2:]), dtype=value.dtype, device=value.device)
            out[cur_key] = torch.cat([pad, value], 1)
            if grad_amplification is not None:
                grad_amplification[i] = list(range(extra))
        elif any(op in key for op in cem_op_names):
            continue                                  # CEM filters are derived, never loaded
        else:
            out[cur_key] = value
    if renamed:
        print('Warning: Modified %d key names due to the change to using ModuleLists' % renamed)
    return out


def load_network(load_path, network, strict=False, optimizer=None, CEM_arch=True, latent_input=None,
                 num_latent_channels=0, scale=4, cem_op_names=None, grad_amplification=None, map_location="cpu"):
    """base_model.py:100-111.  load_path: a file path or an already loaded checkpoint object."""
    if isinstance(network, nn.DataParallel):
        network = network.module
    if isinstance(load_path, os.PathLike):
        load_path = os.fspath(load_path)
    loaded = torch.load(load_path, map_location=map_location) if isinstance(load_path, (str, bytes)) or hasattr(load_path, "read") \
        else load_path
    if 'optimizer_state_dict' in loaded:
        if optimizer is not None:
            optimizer.load_state_dict(loaded['optimizer_state_dict'])
        loaded = loaded['model_state_dict']
    current = network.state_dict()
    if CEM_arch:
        loaded = Adjust_State_Dict_Keys(loaded, current)
    if cem_op_names is None:
        cem_op_names = [n for n, _ in network.named_modules() if 'Filter_OP' in n] if CEM_arch else []
    state = process_loaded_state_dict(loaded, current, latent_input=latent_input, num_latent_channels=num_latent_channels,
                                      scale=scale, cem_op_names=cem_op_names, grad_amplification=grad_amplification)
    network.load_state_dict(state, strict=strict)
    return network


def save_network(save_path, network, optimizer=None):
    """base_model.py:85-97: {'model_state_dict', 'optimizer_state_dict'} with CPU tensors."""
    if isinstance(network, nn.DataParallel):
        network = network.module
    model_state = collections.OrderedDict((k, v.cpu()) for k, v in network.state_dict().items())
    torch.save({'model_state_dict': model_state,
                'optimizer_state_dict': optimizer.state_dict() if optimizer is not None else {}}, save_path)
    return save_path
