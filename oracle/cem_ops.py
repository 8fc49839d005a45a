"""Oracle: CEM operators and projection, fp32 torch-CPU functional ops.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
codes/CEM/CEMnet.py:130-190 (Filter_Layer, CEM_PyTorch.__init__/forward) with the
full 2-D filters exactly as the reference applies them (no separable shortcut).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import cem_filters


class CEMOracle:
    def __init__(self, sf=4, dtype=torch.float32, filters=None, **kw):
        self.sf = sf
        f = filters if filters is not None else cem_filters.derive(sf, **kw)
        self.filters = f
        self.margin_LR, self.margin_HR = f["margin_LR"], f["margin_HR"]
        self.pre, self.post = cem_filters.calc_strides(sf)

        def tile3(k):
            return torch.from_numpy(np.ascontiguousarray(np.tile(k[None, None], (3, 1, 1, 1)))).to(dtype)
        # CEMnet.py:149-159: weights of the three depthwise convs
        self.w_inv = tile3(f["inv_hTh"].astype(np.float32))
        self.w_down = tile3(np.rot90(f["ds_kernel"], 2).astype(np.float32))
        self.w_up = tile3((f["ds_kernel"] * sf ** 2).astype(np.float32))
        self.pad_inv = f["inv_hTh"].shape[0] // 2
        self.pad_aa = f["ds_kernel"].shape[0] // 2

    def _dw(self, x, w, pad):
        return F.conv2d(F.pad(x, (pad,) * 4, mode="replicate"), w.to(x.dtype), groups=3)

    def conv_inv_hTh(self, x):                       # CEMnet.py:149-151
        return self._dw(x, self.w_inv, self.pad_inv)

    def downscale(self, y):                          # CEMnet.py:157-162
        out = self._dw(y, self.w_down, self.pad_aa)
        return out[:, :, self.pre::self.sf, self.pre::self.sf]

    def upscale(self, x):                            # CEMnet.py:153-159
        b, c, h, w = x.shape
        z = x.new_zeros(b, c, h, self.sf, w, self.sf)
        z[:, :, :, self.pre, :, self.pre] = x
        return self._dw(z.view(b, c, h * self.sf, w * self.sf), self.w_up, self.pad_aa)

    def project(self, y, x):                         # CEMnet.py:183-189
        from_lr = self.upscale(self.conv_inv_hTh(x))
        from_gen = self.upscale(self.conv_inv_hTh(self.downscale(y)))
        return from_lr + (y - from_gen)

    # padding of the packed [Z.view, LR] input in eval mode: CEMnet.py:170-181
    def pre_pad(self, model_input, num_latent):
        m, sf = self.margin_LR, self.sf
        if model_input.size(1) == 3 or model_input.size(1) - 3 == num_latent:
            return F.pad(model_input, (m,) * 4, mode="replicate")
        zp, lr = model_input[:, :-3], model_input[:, -3:]
        b, _, h, w = lr.shape
        z = zp.reshape(b, -1, sf * h, sf * w)
        lr = F.pad(lr, (m,) * 4, mode="replicate")
        z = F.pad(z, (sf * m,) * 4, mode="replicate")
        z = z.reshape(b, z.size(1) * sf * sf, lr.size(2), lr.size(3))
        return torch.cat([z, lr], 1)

    def unpad_HR(self, y):                           # CEMnet.py:65
        m = self.margin_HR
        return y[:, :, m:-m, m:-m]


def concat_latent(lr, z, sf=4):
    """SRRaGANModel.ConcatLatent, codes/models/SRRaGAN_model.py:249-255 (raw .view packing)."""
    if z is None:
        return lr.clone()
    if lr.shape[2:] != z.shape[2:]:
        z = z.contiguous().view(z.size(0), z.size(1) * sf * sf, lr.size(2), lr.size(3))
    return torch.cat([z, lr], 1)
