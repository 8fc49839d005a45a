"""Which plans the fused growth launch (ESR_FUSE_RDB=1) handles: python tools/fuse_probe.py tiny_nograd|tiny_grad|mid_nograd|mid_grad
(12 x 14 and 36 x 40 unpadded train-mode plans, with and without a gradient).  engine.GPlan keeps separate launches below 32 x 32."""
import os, sys, torch
sys.path.insert(0, '.')
from esr_b200 import synth
from oracle.cem_ops import concat_latent
from tests.test_gpu_net import build_product_G
dev = torch.device('cuda', 0)
case = sys.argv[1]
wts = synth.make_weights("kaiming", seed=7, nb=2)
h, w, train = (12, 14, True) if case in ('tiny_nograd', 'tiny_grad') else (36, 40, True)
lr, z = synth.make_inputs(1, h, w, seed=7)
netG = build_product_G(dev, 2, "all_layers_HR_downscaled", wts, train=train)
x = concat_latent(lr, z).to(dev)
if case.endswith('nograd'):
    with torch.no_grad():
        out = netG(x)
else:
    x.requires_grad_(True)
    out = netG(x)
torch.cuda.synchronize()
print(case, 'forward ok', float(out.abs().max()))
