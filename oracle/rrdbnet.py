"""Oracle: RRDBNet(+Z) forward and the G+CEM wrapper, fp32 torch-CPU functional ops.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
codes/models/modules/architecture.py:102-175 (RRDBNet), codes/models/modules/block.py
:76-97 (ShortcutBlock), :129-155 (conv_block), :196-242 (ResidualDenseBlock_5C),
:245-270 (RRDB), :294-301 (upconv_blcok) and codes/CEM/CEMnet.py:169-190
(CEM_PyTorch.forward).  Weights are a dict keyed like the reference state_dict
(without the ``generated_image_model.`` prefix).

``operand_dtype=torch.bfloat16`` rounds every conv input and weight to bf16 before
the (fp32) convolution; that emulates the arithmetic of the CUDA path (bf16 MMA
operands, fp32 accumulate, fp32 residual trunk) so that tests can separate
indexing errors from rounding.
"""
import math

import torch
import torch.nn.functional as F

from .cem_ops import CEMOracle


def _q(t, operand_dtype):
    return t if operand_dtype is None else t.to(operand_dtype).to(torch.float32)


class RRDBNetOracle:
    def __init__(self, weights, nb=23, upscale=4, latent_input="all_layers_HR_downscaled",
                 num_latent_channels=3, operand_dtype=None):
        self.w = weights
        self.nb, self.upscale = nb, upscale
        self.latent_input = latent_input
        self.nz = num_latent_channels if latent_input is not None else 0
        self.all_layers = latent_input is not None and "all_layers" in latent_input
        self.n_up = 1 if upscale == 3 else int(math.log(upscale, 2))
        self.od = operand_dtype

    def conv(self, x, key, act):                     # block.py:129-155 (+ act :10-23)
        y = F.conv2d(_q(x, self.od), _q(self.w[key + ".weight"], self.od), self.w[key + ".bias"], padding=1)
        return F.leaky_relu(y, 0.2) if act else y

    def rdb(self, x, prefix):                        # block.py:230-235
        outs = [x]
        for i in range(5):
            outs.append(self.conv(torch.cat(outs, 1), "%s.convs.%d.0" % (prefix, i), act=i < 4))
        return outs[-1] * 0.2 + x[:, -outs[-1].size(1):]

    def rrdb(self, x, prefix, lat):                  # block.py:262-270
        out = self.rdb(x, prefix + ".RDB1")
        for name in ("RDB2", "RDB3"):
            if lat is not None:
                out = torch.cat([lat, out], 1)
            out = self.rdb(out, "%s.%s" % (prefix, name))
        return out * 0.2 + x[:, -out.size(1):]

    def trunk(self, x, lat):                         # block.py:85-97 (model.1)
        out = x
        for r in range(self.nb):
            if r > 0 and lat is not None:
                out = torch.cat([lat, out], 1)
            out = self.rrdb(out, "model.1.sub.%d" % r, lat)
        if lat is not None:
            out = torch.cat([lat, out], 1)
        out = self.conv(out, "model.1.sub.%d" % self.nb, act=False)
        nz = lat.size(1) if lat is not None else 0
        return x[:, nz:] + out

    def forward(self, x):                            # architecture.py:151-175
        lat_lr = lat_hr = None
        if self.latent_input is not None:
            assert "HR_downscaled" in self.latent_input, "oracle covers the HR_downscaled domain"
            zp, x = x[:, :-3], x[:, -3:]
            b, _, h, w = x.shape
            lat_hr = zp.reshape(b, -1, self.upscale * h, self.upscale * w)
            lat_lr = F.interpolate(lat_hr, scale_factor=1 / self.upscale, mode="bilinear", align_corners=False)
            x = torch.cat([lat_lr, x], 1)
        x = self.conv(x, "model.0", act=False)
        lat = lat_lr if self.all_layers else None
        if lat is not None:
            x = torch.cat([lat, x], 1)
        x = self.trunk(x, lat)
        for u in range(self.n_up):                   # block.py:294-301, no latent (architecture.py:164-171)
            x = F.interpolate(x, scale_factor=3 if self.upscale == 3 else 2, mode="nearest")
            x = self.conv(x, "model.%d.1" % (2 + u), act=True)
        lat = lat_hr if self.all_layers else None
        k0 = 2 + self.n_up
        x = self.conv(torch.cat([lat, x], 1) if lat is not None else x, "model.%d" % k0, act=True)
        x = self.conv(torch.cat([lat, x], 1) if lat is not None else x, "model.%d" % (k0 + 2), act=False)
        return x


class GCEMOracle:
    """CEM_PyTorch wrapping RRDBNet: CEMnet.py:142-194."""

    def __init__(self, weights, sf=4, pre_pad=True, cem=None, **net_kw):
        self.net = RRDBNetOracle(weights, upscale=sf, **net_kw)
        self.cem = cem if cem is not None else CEMOracle(sf)
        self.pre_pad = pre_pad

    def forward(self, model_input, return_raw=False):
        x = self.cem.pre_pad(model_input, self.net.nz) if self.pre_pad else model_input
        y = self.net.forward(x)
        out = self.cem.project(y, x[:, -3:])
        out = self.cem.unpad_HR(out) if self.pre_pad else out
        return (out, y) if return_raw else out
