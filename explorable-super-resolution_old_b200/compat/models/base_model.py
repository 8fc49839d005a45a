"""`from models.base_model import BaseModel` -> checkpoint import / export of the B200 generator
(reference: codes/models/base_model.py:85-144).  Only the loader / saver methods exist; the training plumbing of the
reference's BaseModel (schedulers, logging, receptive-field helpers) is outside the built hot path."""
from esr_b200 import checkpoint as _ckpt


class BaseModel:
    def __init__(self, opt):
        self.opt = opt
        self.channels_idx_4_grad_amplification = {}

    def process_loaded_state_dict(self, loaded_state_dict, current_state_dict):
        names = getattr(getattr(self, 'CEM_net', None), 'OP_names', []) if getattr(self, 'CEM_arch', False) else []
        return _ckpt.process_loaded_state_dict(
            loaded_state_dict, current_state_dict, latent_input=getattr(self, 'latent_input', None),
            num_latent_channels=getattr(self, 'num_latent_channels', 0) or 0, scale=self.opt['scale'], cem_op_names=names,
            grad_amplification=self.channels_idx_4_grad_amplification)

    def load_network(self, load_path, network, strict=False, optimizer=None):
        names = getattr(getattr(self, 'CEM_net', None), 'OP_names', None) if self.opt['network_G']['CEM_arch'] else []
        return _ckpt.load_network(load_path, network, strict=strict, optimizer=optimizer,
                                  CEM_arch=bool(self.opt['network_G']['CEM_arch']),
                                  latent_input=getattr(self, 'latent_input', None),
                                  num_latent_channels=getattr(self, 'num_latent_channels', 0) or 0, scale=self.opt['scale'],
                                  cem_op_names=names, grad_amplification=self.channels_idx_4_grad_amplification)

    def save_network(self, save_dir, network, network_label, iter_label, optimizer):
        import os
        return _ckpt.save_network(os.path.join(save_dir, '{}_{}.pth'.format(iter_label, network_label)), network, optimizer)
