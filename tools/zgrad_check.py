"""Gradient parity (relative error / cosine vs the CPU oracle's autograd) and speed of the data-gradient backward, for
A/B of backward operand formats (ESR_BWD_PRECISE=0|1).  nb=23 at 1x3x16x16 for parity, config 3 for speed."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esr_b200 import synth  # noqa: E402
from oracle.cem_ops import concat_latent  # noqa: E402
from oracle.rrdbnet import GCEMOracle  # noqa: E402
from tests.test_gpu_net import build_product_G  # noqa: E402

dev = torch.device("cuda", 0)
for kind, seed, nb, hw in (("default", 1, 23, 16), ("kaiming", 7, 2, 14)):
    wts = synth.make_weights(kind, seed=seed, nb=nb)
    lr, z = synth.make_inputs(1, hw, hw, seed=seed)
    gout = torch.randn(1, 3, 4 * hw, 4 * hw, generator=torch.Generator().manual_seed(5))
    zo = z.clone().requires_grad_(True)
    (GCEMOracle(wts, nb=nb).forward(concat_latent(lr, zo)) * gout).sum().backward()
    netG = build_product_G(dev, nb, "all_layers_HR_downscaled", wts)
    zp = z.clone().to(dev).requires_grad_(True)
    (netG(concat_latent(lr.to(dev), zp)) * gout.to(dev)).sum().backward()
    got, ref = zp.grad.cpu(), zo.grad
    print("%s nb=%d: rel %.4f cos %.6f" % (kind, nb, float((got - ref).norm() / ref.norm()),
                                          float((got * ref).sum() / (got.norm() * ref.norm()))), flush=True)
wts = synth.make_weights("default", seed=0)
netG = build_product_G(dev, 23, "all_layers_HR_downscaled", wts)
netG.generated_image_model.use_cuda_graphs = True
lr, z = synth.make_inputs(1, 256, 256, seed=3)
lr = lr.to(dev)
zp = z.to(dev).requires_grad_(True)
for _ in range(3):
    out = netG(concat_latent(lr, zp)); out.abs().mean().backward(); zp.grad = None
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = netG(concat_latent(lr, zp)); out.abs().mean().backward(); zp.grad = None
e1.record(); torch.cuda.synchronize()
print("config 3 forward+backward (graphs): %.2f ms" % (e0.elapsed_time(e1) / 5))
