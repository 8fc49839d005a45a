"""Editing by optimising the control signal Z through the frozen G+CEM.  Drop-in for the hot-loop part
of the reference's ``Z_optimization.py``: ``Z_optimizer`` (:326-682, loop :555-655), ``Optimizable_Z``
(:272-317), ``ArcTanH`` (:319-320), ``TV_Loss`` (:322-324), same constructor arguments and attributes.

The generator forward and its data-gradient backward run in libesr_b200.so; the scalar objectives are
plain torch ops on ``fake_H`` (they only consume the hot path's output and hand back dL/d fake_H).
Objectives: 'l1' (training mode, HR_unpadder given), 'TV', 'max_STD' / 'min_STD' / 'STD_increase' /
'STD_decrease' (global; these run as one CUDA graph per iteration when nothing else is in the loop), and -
through ``z_objectives`` - the GUI's 'local_' STD / Mag variants, histogram and
dictionary imitation (the density sums are kernels of libesr_b200.so), periodicity, scribbles and the
diverse-solutions 'random_l1' objectives.  Not built: VGG, adversarial, desired_SVD, automatic temperature.
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
_FUSED = ('TV', 'max_STD', 'min_STD', 'STD_increase', 'STD_decrease')
HIST_LEN = 4096


class _FusedZLoop:
    """The whole iteration of Z_optimizer.optimize (Z_optimization.py:572-635) as ONE CUDA graph of this package's
    kernels: tanh + packing -> G forward -> CEM -> objective -> its gradient -> CEM adjoint -> G data-gradient ->
    tanh' + Adam.  No ATen kernel and no host read inside the loop; losses are collected on the device and read once."""

    def __init__(self, zopt, wrapper, G, lr_img, key):
        import ctypes as C
        from . import _capi as capi
        from .rrdbnet import _forward_eager, _capture
        from .backward import generator_backward_eager
        self.key = key
        Zp = zopt.Z_model.Z
        dev = Zp.device
        # the graph works on its own copy of Z: Optimizable_Z.forward re-assigns Z.data (its clamp), so the parameter's
        # storage is not a stable address across optimize() calls; run() copies Z in before and out after the replays
        self.zbuf = Zp.data.clone()
        B, nz = Zp.shape[0], Zp.shape[1]
        h, w = lr_img.shape[2:]
        sf = G.upscale
        margin = wrapper._margin_LR if wrapper.pre_pad else 0
        filters = wrapper._filters
        self.B, self.n_lat, self.n_img = B, nz * sf * sf * h * w, (nz * sf * sf + 3) * h * w
        self.x = torch.empty(B, nz * sf * sf + 3, h, w, device=dev, dtype=torch.float32)
        self.x[:, nz * sf * sf:] = lr_img.to(dev).float().expand(B, -1, -1, -1)
        plan = G.plan(B, h, w, margin, keep=True)
        bp = G.backward_plan(plan)
        crop = sf * margin
        H, W = sf * plan.hp - 2 * crop, sf * plan.wp - 2 * crop
        onc = plan.y.size(1)
        self.out = torch.empty(B, onc, H, W, device=dev, dtype=torch.float32)
        self.ws = torch.empty(max(1, 2 * B * onc * plan.hp * plan.wp), device=dev, dtype=torch.float32)
        self.g = torch.zeros_like(self.out)
        l = capi.lib()
        self.red = torch.empty(int(l.esr_zopt_loss_workspace_floats(B, H)), device=dev, dtype=torch.float32)
        self.stats = torch.zeros(B, 8, device=dev, dtype=torch.float32)
        self.hist = torch.zeros(HIST_LEN, device=dev, dtype=torch.float32)
        self.step = torch.zeros(1, device=dev, dtype=torch.int32)
        opt = zopt.optimizer
        grp = opt.param_groups[0]
        st = opt.state[Zp]                           # created by Z_optimizer._fused_loop when torch had not yet
        self.state = st
        mode, sign, w_std, target = zopt._fused_objective()
        self.target = target.to(dev).float().reshape(-1).expand(B).contiguous() if target is not None else None
        self.plan, self.bp, self.keep = plan, bp, (wrapper, G, filters)
        z_range = float(zopt.Z_model.Z_range)
        lr, (b1, b2), eps = float(grp['lr']), grp['betas'], float(grp['eps'])

        def iteration():
            sp = capi.stream_ptr()
            capi.check(l.esr_zopt_tanh_pack(capi.ptr(self.zbuf), z_range, B, self.n_lat, self.n_img, capi.ptr(self.x), sp))
            _forward_eager(plan, self.x, filters, crop, self.out, self.ws)
            capi.check(l.esr_zopt_loss(capi.ptr(self.out), B, onc, H, W, mode, sign, w_std, capi.ptr(self.target), capi.ptr(self.red),
                                       capi.ptr(self.stats), capi.ptr(self.hist), HIST_LEN, capi.ptr(self.step), sp))
            capi.check(l.esr_zopt_loss_grad(capi.ptr(self.out), B, onc, H, W, capi.ptr(self.stats), capi.ptr(self.g), sp))
            g_in = generator_backward_eager(plan, bp, filters, margin, self.g)
            capi.check(l.esr_zopt_adam(capi.ptr(self.zbuf), capi.ptr(st['exp_avg']), capi.ptr(st['exp_avg_sq']), capi.ptr(g_in), z_range, B,
                                       self.n_lat, self.n_img, lr, float(b1), float(b2), eps, capi.ptr(self.step), sp))
            return g_in
        # the warm-up pass that precedes the capture is a real iteration: put the state back afterwards
        keep = [t.clone() for t in (st['exp_avg'], st['exp_avg_sq'])]
        with torch.cuda.device(dev):
            self.graph, self.g_in = _capture(iteration, dev)
        for t, k in zip((st['exp_avg'], st['exp_avg_sq']), keep):
            t.copy_(k)
        self.launches = None

    def run(self, iters, Zp):
        """Replays `iters` iterations on the parameter Zp; returns (loss per iteration, per-image losses of the last one)."""
        k0 = int(float(self.state['step']))
        self.step.fill_(k0)
        self.zbuf.copy_(Zp.data)
        for _ in range(iters):
            self.graph.replay()
        Zp.data.copy_(self.zbuf)
        idx = (torch.arange(k0, k0 + iters, device=self.hist.device) % HIST_LEN)
        losses = self.hist[idx].cpu().tolist()                     # the one host read of the loop
        latest = self.stats[:, 4].cpu().tolist()
        self.state['step'] = torch.tensor(float(k0 + iters))
        return losses, latest


class Z_optimizer():
    MIN_LR = 1e-5
    PATCH_SIZE_4_STD = 7

    def __init__(self, objective, Z_size, model, Z_range, max_iters, data=None, loggers=None, image_mask=None, Z_mask=None,
                 initial_Z=None, initial_LR=None, existing_optimizer=None, batch_size=1, HR_unpadder=None,
                 auto_set_hist_temperature=False, random_Z_inits=False):
        from . import z_objectives
        if objective not in _BUILT:
            why = z_objectives.unsupported_reason(objective, auto_set_hist_temperature)
            if why is not None:
                raise NotImplementedError("Z objective %r %s (built into the single-graph loop: %s)" % (objective, why, ', '.join(_BUILT)))
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
                                     initial_pre_tanh_Z=initial_pre_tanh_Z, Z_mask=Z_mask, device=self.device,
                                     random_perturbations=(random_Z_inits and 'random' not in objective) or
                                     ('random' in objective and 'limited' in objective))
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
        # Masked_STD: global, or per 7x7 patch for the 'local' objectives (Z_optimization.py:356-361, :525-535)
        self._std = z_objectives.LocalStd(self, 'local' in objective, image_mask, self.PATCH_SIZE_4_STD)
        self._rich = None
        if not self.model_training:
            self.initial_STD = self.Masked_STD(first_image_only=True)
            print('Initial STD: %.3e' % (self.initial_STD.mean().item()))
        if existing_optimizer is None:
            if any(p in objective for p in ('l1', 'scribble')) and 'random' not in objective:
                if data is not None and 'HR' in data.keys():
                    self.GT_HR = data['HR']
                if self.image_mask is None:
                    self.loss = torch.nn.L1Loss().to(self.device)
                elif 'scribble' in objective:
                    self._rich = self.loss = z_objectives.ScribbleObjective(self, data)
                else:
                    raise NotImplementedError("'l1' with an image mask only exists as the scribble objective (the reference's "
                                              "masked l1 reads masks that only the scribble set-up defines, Z_optimization.py:390-398)")
            elif objective not in _BUILT:
                self._rich = z_objectives.resolve(self, data, auto_set_hist_temperature)
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

    def _fused_objective(self):
        """(mode, sign, std weight, per-image target STD) of esr_zopt_loss for this objective."""
        o = self.objective
        if o == 'TV':
            return 0, 1.0, float(self.STD_PRESERVING_WEIGHT), self.initial_STD
        if o in ('max_STD', 'min_STD'):
            return 1, (-1.0 if 'max' in o else 1.0), 0.0, None
        return 2, 1.0, 0.0, self.desired_STD

    def _fused_loop(self):
        """The single-graph loop when everything in the iteration is this package's: a built objective without masks, a
        fixed iteration count, no loggers, GUI (eval) mode, torch's plain Adam on Z alone, and a netG that is this
        package's CEM-wrapped RRDBNet with a latent input.  None otherwise (the generic loop below runs instead)."""
        from .cem import CEM_PyTorch
        from .rrdbnet import RRDBNet
        if os.environ.get('ESR_ZOPT_FUSED', '1') == '0' or self.objective not in _FUSED:
            return None
        if not (0 < self.max_iters <= HIST_LEN) or self.loggers is not None or self.model_training or self.scheduler is not None:
            return None
        if self.Z_model.mask is not None or self.Z_mask is not None or self.Z_model.Z_range is None:
            return None
        if self.objective in ('STD_increase', 'STD_decrease', 'TV') and not hasattr(self, 'desired_STD' if 'STD' in self.objective else 'STD_PRESERVING_WEIGHT'):
            return None
        wrapper = getattr(self.model.netG, 'module', self.model.netG)
        G = getattr(wrapper, 'generated_image_model', None)
        if not isinstance(wrapper, CEM_PyTorch) or not isinstance(G, RRDBNet) or G._cfg['nz_in'] == 0:
            return None
        opt = self.optimizer
        if type(opt) is not torch.optim.Adam or len(opt.param_groups) != 1:
            return None
        grp = opt.param_groups[0]
        if len(grp['params']) != 1 or grp['params'][0] is not self.Z_model.Z or grp.get('amsgrad') or grp.get('weight_decay') \
                or grp.get('maximize') or grp.get('capturable') or grp.get('differentiable'):
            return None
        Zp, lr_img = self.Z_model.Z, self.data['LR']
        sf = G.upscale
        if not Zp.is_cuda or Zp.dtype != torch.float32 or not Zp.is_contiguous() or lr_img.dim() != 4 or \
                lr_img.size(0) not in (1, Zp.size(0)) or tuple(Zp.shape[2:]) != (sf * lr_img.size(2), sf * lr_img.size(3)) or \
                Zp.size(1) != G._cfg['nz_in'] or (lr_img.size(2) * lr_img.size(3)) % 4 != 0:
            return None
        if self.objective in ('TV', 'STD_increase', 'STD_decrease'):
            tgt = self.initial_STD if self.objective == 'TV' else self.desired_STD
            if tgt.numel() not in (1, Zp.size(0)):
                return None
        st = opt.state[Zp]
        if 'exp_avg' not in st:                     # torch.optim.Adam creates these lazily at its first step()
            st['step'] = torch.tensor(0.0)
            st['exp_avg'] = torch.zeros_like(Zp, memory_format=torch.preserve_format)
            st['exp_avg_sq'] = torch.zeros_like(Zp, memory_format=torch.preserve_format)
        key = (tuple(Zp.shape), tuple(lr_img.shape), wrapper.pre_pad, id(wrapper._filters), float(grp['lr']), tuple(grp['betas']),
               float(grp['eps']), self.objective, tuple((p.data_ptr(), p._version) for p in G.parameters()),
               float(self.Z_model.Z_range), lr_img.data_ptr(), lr_img._version,
               st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr())
        cur = getattr(self, '_fused', None)
        if cur is None or cur.key != key:
            if cur is not None and os.environ.get('ESR_ZOPT_DEBUG'):
                import sys
                print('fused Z loop re-captured; key fields that changed:', [(i, a, b) for i, (a, b) in enumerate(zip(cur.key, key)) if a != b and i != 8], file=sys.stderr)
            self._fused = None                      # free the old graph's buffers first
            self._fused = _FusedZLoop(self, wrapper, G, lr_img, key)
        return self._fused

    def Masked_STD(self, first_image_only=False):
        return self._std(first_image_only)

    def feed_data(self, data):
        self.data = data
        self.cur_iter = 0
        if 'l1' in self.objective:
            self.GT_HR = data['HR'].to(self.device)
        elif 'hist' in self.objective:
            self.loss.Feed_Desired_Hist_Im(data['HR'].to(self.device))

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
        fused = self._fused_loop()
        while fused is not None:                    # one pass: every iteration is a replay of one CUDA graph
            self.loss_values, self.latest_Z_loss_values = fused.run(self.max_iters, self.Z_model.Z)
            z_iter += self.max_iters
            self.data['Z'] = fused.x[:, :-3].reshape(self.Z_model.Z.shape).clone()     # the Z of the last forward, as in the reference
            self.model.feed_data(self.data, need_HR=False)
            self.model.fake_H = fused.out.clone()
            defer, pending_latest = False, None
            break
        while fused is None:
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
            if self._rich is not None:
                Z_loss = self._rich(fake_H)
            elif 'l1' in self.objective:
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
            if Z_loss.dim() == 0:            # plain 'hist': one KL divergence for the batch (the reference's loop cannot
                Z_loss = Z_loss.reshape(1)   # iterate over it, Z_optimization.py:622)
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
        if 'random' in self.objective and 'limited' in self.objective and len(self.loss_values) > 1:
            self.loss_values[0] = self.loss_values[1]          # the first value is ~0 there (Z_optimization.py:638-639)
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
