"""Editing by optimising the control signal Z through the frozen G+CEM.  Drop-in for the hot-loop part
of the reference's ``Z_optimization.py``: ``Z_optimizer`` (:326-682, loop :555-655), ``Optimizable_Z``
(:272-317), ``ArcTanH`` (:319-320), ``TV_Loss`` (:322-324), same constructor arguments and attributes.

The generator forward and its data-gradient backward run in libesr_b200.so; the scalar objectives are
plain torch ops on ``fake_H`` (they only consume the hot path's output and hand back dL/d fake_H).
Objectives built here: 'l1' (training mode, HR_unpadder given), 'TV', 'max_STD' / 'min_STD' /
'STD_increase' / 'STD_decrease' (global).  The histogram / dictionary / scribble / periodicity / VGG /
adversarial objectives of the GUI are out of this path's scope.
"""
import os
import time

import numpy as np
import torch


def ArcTanH(input_tensor):
    eps = torch.finfo(input_tensor.dtype).eps
    return 0.5 * torch.log((1 + input_tensor + eps) / (1 - input_tensor + eps))


def TV_Loss(image):
    return (image[:, :, :, :-1] - image[:, :, :, 1:]).abs().mean(dim=(1, 2, 3)) + \
           (image[:, :, :-1, :] - image[:, :, 1:, :]).abs().mean(dim=(1, 2, 3))


class Optimizable_Z(torch.nn.Module):
    def __init__(self, Z_shape, Z_range=None, initial_pre_tanh_Z=None, Z_mask=None, random_perturbations=False,
                 device=None):
        super().__init__()
        device = device or torch.device('cuda')
        self.Z = torch.nn.Parameter(torch.zeros(Z_shape, dtype=torch.float32, device=device))
        self.mask = None
        if Z_mask is not None and not np.all(Z_mask):
            self.mask = torch.from_numpy(Z_mask).float().to(device)
            self.initial_pre_tanh_Z = 1 * initial_pre_tanh_Z.float().to(device)
        if initial_pre_tanh_Z is not None:
            assert initial_pre_tanh_Z.size()[1:] == self.Z.data.size()[1:] and \
                initial_pre_tanh_Z.size(0) in [1, self.Z.data.size(0)], 'Initilizer size does not match desired Z size'
            if random_perturbations:
                initial_pre_tanh_Z = initial_pre_tanh_Z + 0.001 * torch.randn_like(initial_pre_tanh_Z)
            self.Z.data[:initial_pre_tanh_Z.size(0), ...] = initial_pre_tanh_Z.to(device)
        self.Z_range = Z_range
        if Z_range is not None:
            self.tanh = torch.nn.Tanh()

    def forward(self):
        if self.Z_range is not None:
            big = torch.finfo(self.Z.dtype).max
            self.Z.data = self.Z.data.clamp(-big, big)
        if self.mask is not None:
            self.Z.data = self.mask * self.Z.data + (1 - self.mask) * self.initial_pre_tanh_Z
        return self.Z_range * self.tanh(self.Z) if self.Z_range is not None else self.Z

    def PreTanhZ(self):
        if self.mask is not None:
            return self.mask * self.Z.data + (1 - self.mask) * self.initial_pre_tanh_Z
        return self.Z.data

    def Randomize_Z(self, what_2_shuffle):
        assert what_2_shuffle in ['all', 'allButFirst']
        if what_2_shuffle == 'all':
            torch.nn.init.xavier_uniform_(self.Z.data, gain=100)
        else:
            torch.nn.init.xavier_uniform_(self.Z.data[1:], gain=100)

    def Return_Detached_Z(self):
        return self.forward().detach()


_BUILT = ('l1', 'TV', 'max_STD', 'min_STD', 'STD_increase', 'STD_decrease')


class Z_optimizer():
    MIN_LR = 1e-5
    PATCH_SIZE_4_STD = 7

    def __init__(self, objective, Z_size, model, Z_range, max_iters, data=None, loggers=None, image_mask=None, Z_mask=None,
                 initial_Z=None, initial_LR=None, existing_optimizer=None, batch_size=1, HR_unpadder=None,
                 auto_set_hist_temperature=False, random_Z_inits=False):
        if objective not in _BUILT:
            raise NotImplementedError("Z objective %r is outside the built hot path (built: %s)" % (objective, ', '.join(_BUILT)))
        self.device = next(model.netG.parameters()).device
        if initial_Z is not None or 'cur_Z' in model.__dict__.keys():
            if initial_Z is None:
                initial_Z = 1 * model.GetLatent()
            initial_pre_tanh_Z = initial_Z / Z_range
            eps = torch.finfo(initial_pre_tanh_Z.dtype).eps
            initial_pre_tanh_Z = ArcTanH(torch.clamp(initial_pre_tanh_Z, min=-1 + eps, max=1. - eps))
        else:
            initial_pre_tanh_Z = None
        self.Z_model = Optimizable_Z(Z_shape=[batch_size, model.num_latent_channels] + list(Z_size), Z_range=Z_range,
                                     initial_pre_tanh_Z=initial_pre_tanh_Z, Z_mask=Z_mask, random_perturbations=random_Z_inits,
                                     device=self.device)
        assert (initial_LR is not None) or (existing_optimizer is not None), \
            'Should either supply optimizer from previous iterations or initial LR for new optimizer'
        self.objective, self.data, self.model = objective, data, model
        self.model_training = HR_unpadder is not None
        if image_mask is None:
            self.image_mask = torch.ones(list(model.fake_H.size()[2:]), dtype=model.fake_H.dtype, device=self.device) \
                if 'fake_H' in model.__dict__.keys() else None
            self.Z_mask = None
        else:
            assert Z_mask is not None, 'Should either supply both masks or niether'
            self.image_mask = torch.from_numpy(image_mask).to(model.fake_H.dtype).to(self.device)
            self.Z_mask = torch.from_numpy(Z_mask).to(model.fake_H.dtype).to(self.device)
            self.initial_Z = 1. * model.GetLatent()
        if not self.model_training:
            self.initial_STD = self.Masked_STD(first_image_only=True)
            print('Initial STD: %.3e' % (self.initial_STD.mean().item()))
        if existing_optimizer is None:
            if objective == 'l1':
                if data is not None and 'HR' in data.keys():
                    self.GT_HR = data['HR']
                if self.image_mask is not None:
                    raise NotImplementedError("'l1' with an image mask is the scribble objective (not built)")
                self.loss = torch.nn.L1Loss().to(self.device)
            elif 'STD' in objective:
                if any(p in objective for p in ['increase', 'decrease']):
                    inc = data['STD_increment']
                    self.desired_STD = self.initial_STD
                    if inc is None:
                        self.desired_STD *= 1.05 if 'increase' in objective else 1 / 1.05
                    else:
                        self.desired_STD += inc if 'increase' in objective else -inc
            elif objective == 'TV':
                self.STD_PRESERVING_WEIGHT = 100
            self.optimizer = torch.optim.Adam(self.Z_model.parameters(), lr=initial_LR)
        else:
            self.optimizer = existing_optimizer
        self.LR = initial_LR
        self.scheduler = None
        self.loggers = loggers
        self.cur_iter = 0
        self.max_iters = max_iters
        self.random_Z_inits = 'all' if (random_Z_inits or self.model_training) \
            else 'allButFirst' if (initial_pre_tanh_Z is not None and initial_pre_tanh_Z.size(0) < batch_size) else False
        self.HR_unpadder = HR_unpadder

    def Masked_STD(self, first_image_only=False):
        return torch.std(self.model.fake_H * self.image_mask, dim=(1, 2, 3)).view(1, -1)

    def feed_data(self, data):
        self.data = data
        self.cur_iter = 0
        if 'l1' in self.objective:
            self.GT_HR = data['HR'].to(self.device)

    def Manage_Model_Grad_Requirements(self, disable):
        if disable:
            self.original_requires_grad_status = []
            for p in self.model.netG.parameters():
                self.original_requires_grad_status.append(p.requires_grad)
                p.requires_grad = False
        else:
            for i, p in enumerate(self.model.netG.parameters()):
                p.requires_grad = self.original_requires_grad_status[i]

    def optimize(self):
        self.Manage_Model_Grad_Requirements(disable=True)
        G = getattr(self.model.netG, 'generated_image_model', self.model.netG)
        if hasattr(G, 'use_cuda_graphs') and os.environ.get('ESR_ZOPT_GRAPH', '1') != '0':
            G.use_cuda_graphs = True           # forward and backward of every iteration replay as CUDA graphs
        # With a fixed iteration count nothing inside the loop needs a loss value on the host, so the per-iteration
        # .item() reads of the reference loop (Z_optimization.py:620-634) are deferred to one read after the last
        # iteration: the host queues iteration k+1 while the GPU runs iteration k.  loss_values / latest_Z_loss_values
        # hold plain floats when optimize() returns, as in the reference.  The convergence mode (max_iters < 0) and
        # loggers read every value as they go.
        defer = self.max_iters > 0 and self.loggers is None and os.environ.get('ESR_ZOPT_DEFER_READS', '1') != '0'
        pending_latest = None
        self.loss_values = []
        if self.random_Z_inits and self.cur_iter == 0:
            self.Z_model.Randomize_Z(what_2_shuffle=self.random_Z_inits)
        z_iter = self.cur_iter
        while True:
            if self.max_iters > 0:
                if z_iter == (self.cur_iter + self.max_iters):
                    break
            elif len(self.loss_values) >= -self.max_iters:
                if z_iter == (self.cur_iter - 5 * self.max_iters):
                    break
                if (self.loss_values[self.max_iters] - self.loss_values[-1]) / np.abs(self.loss_values[self.max_iters]) < 1e-2 * self.LR:
                    break
            self.optimizer.zero_grad()
            self.data['Z'] = self.Z_model()
            self.model.feed_data(self.data, need_HR=False)
            self.model.fake_H = self.model.netG(self.model.model_input)      # G+CEM forward, autograd on
            if self.model_training:
                self.model.fake_H = self.HR_unpadder(self.model.fake_H)
            fake_H = self.model.fake_H
            if self.objective == 'l1':
                Z_loss = self.loss(fake_H, self.GT_HR.to(self.device))
            elif 'STD' in self.objective:
                Z_loss = self.Masked_STD(first_image_only=False)
                if any(p in self.objective for p in ['increase', 'decrease']):
                    Z_loss = (Z_loss - self.desired_STD) ** 2
                Z_loss = Z_loss.mean(0)
            else:  # 'TV'
                Z_loss = (self.STD_PRESERVING_WEIGHT * (self.Masked_STD(first_image_only=False) - self.initial_STD) ** 2).mean(0) + \
                    TV_Loss(fake_H * self.image_mask)
            if 'max' in self.objective:
                Z_loss = -1 * Z_loss
            cur_LR = self.optimizer.param_groups[0]['lr']
            if self.loggers is not None:
                for logger_num, logger in enumerate(self.loggers):
                    cur_value = Z_loss[logger_num].item() if Z_loss.dim() > 0 else Z_loss.item()
                    logger.print_format_results('val', {'epoch': 0, 'iters': z_iter, 'time': time.time(), 'model': '',
                                                        'lr': cur_LR, 'Z_loss': cur_value}, dont_print=True)
            if not self.model_training:
                if defer:
                    pending_latest = Z_loss.detach()
                else:
                    self.latest_Z_loss_values = [val.item() for val in Z_loss]
            Z_loss = Z_loss.mean()
            Z_loss.backward()                                                # data gradient back to Z
            self.loss_values.append(Z_loss.detach() if defer else Z_loss.item())
            self.optimizer.step()
            z_iter += 1
        if defer:
            if self.loss_values:
                self.loss_values = torch.stack(self.loss_values).cpu().tolist()
            if pending_latest is not None:
                self.latest_Z_loss_values = pending_latest.cpu().tolist()
        if not self.model_training:
            print('Final STDs: ', ['%.3e' % (val.item()) for val in self.Masked_STD(first_image_only=False).mean(0)])
        self.cur_iter = z_iter + 1
        Z_2_return = self.Z_model.Return_Detached_Z()
        self.Manage_Model_Grad_Requirements(disable=False)
        if self.model_training:
            # one more un-cropped forward for the caller; the generator's weight gradients are not built,
            # so it runs without autograd (the reference runs it with, for the GAN training step)
            self.data['Z'] = Z_2_return
            self.model.feed_data(self.data, need_HR=False)
            with torch.no_grad():
                self.model.fake_H = self.model.netG(self.model.model_input)
        return Z_2_return

    def ReturnStatus(self):
        return self.Z_model.PreTanhZ(), self.optimizer


class SRModelShim:
    """The slice of ``SRRaGANModel`` (codes/models/SRRaGAN_model.py:249-302, :577-584) that the hot path's
    callers touch: ConcatLatent / GetLatent / feed_data / test around a ``netG``."""

    def __init__(self, netG, scale=4, num_latent_channels=3):
        self.netG, self.opt, self.num_latent_channels = netG, {'scale': scale}, num_latent_channels
        self.device = next(netG.parameters()).device

    def ConcatLatent(self, LR_image, latent_input):
        if latent_input is not None:
            if LR_image.size()[2:] != latent_input.size()[2:]:
                s = self.opt['scale']
                latent_input = latent_input.contiguous().view([latent_input.size(0), latent_input.size(1) * s ** 2] +
                                                              list(LR_image.size()[2:]))
            self.model_input = torch.cat([latent_input, LR_image], dim=1)
        else:
            self.model_input = 1 * LR_image

    def GetLatent(self):
        latent = 1 * self.model_input[:, :-3, ...]
        if latent.size(1) != self.num_latent_channels:
            s = self.opt['scale']
            latent = latent.view([latent.size(0), self.num_latent_channels] + [s * v for v in latent.size()[2:]])
        return latent

    def feed_data(self, data, need_HR=True):
        self.var_L = data['LR'].to(self.device)
        cur_Z = data['Z'] if 'Z' in data.keys() else \
            2 * torch.rand([self.var_L.size(0), self.num_latent_channels, 1, 1]) - 1
        if cur_Z.size(2) == 1:
            s = self.opt['scale']
            cur_Z = (cur_Z * torch.ones([1, 1] + [s * v for v in self.var_L.size()[2:]])).type(self.var_L.type())
        self.ConcatLatent(LR_image=self.var_L, latent_input=cur_Z.to(self.device))
        if need_HR and 'HR' in data:
            self.var_H = data['HR'].to(self.device)

    def test(self, prevent_grads_calc=True):
        self.netG.eval()
        if prevent_grads_calc:
            with torch.no_grad():
                self.fake_H = self.netG(self.model_input)
        else:
            self.fake_H = self.netG(self.model_input)
        self.netG.train()
