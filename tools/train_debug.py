"""Bisects the generator training step at the config-5 shape (16 x 3x32x32, nb=23, train mode): forward, data gradients, weight gradients."""
import os, sys, torch
sys.path.insert(0, '.')
os.environ.setdefault('CUDA_LAUNCH_BLOCKING', '1')
from esr_b200 import synth
from esr_b200.training import GeneratorTrainer
from oracle.cem_ops import concat_latent
from tests.test_gpu_net import build_product_G
dev = torch.device('cuda', 0)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 23
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
h = int(sys.argv[3]) if len(sys.argv) > 3 else 32
wts = synth.make_weights("kaiming", seed=0, nb=nb)
netG = build_product_G(dev, nb, "all_layers_HR_downscaled", wts, train=True)
lr, z = synth.make_inputs(B, h, h, seed=0)
mi = concat_latent(lr, z).to(dev)
tr = GeneratorTrainer(netG)
fake = tr.forward(mi); torch.cuda.synchronize(); print("forward ok", float(fake.abs().max()))
g = torch.randn_like(fake)
plan, filters, margin = tr._state
from esr_b200.backward import generator_backward_eager
bp = tr._backward_plan(plan); torch.cuda.synchronize(); print("backward plan ok")
gi = generator_backward_eager(plan, bp, filters, margin, g); torch.cuda.synchronize(); print("dgrad ok", float(gi.abs().max()))
tr._state = (plan, filters, margin)
tr.backward(g); torch.cuda.synchronize(); print("wgrad ok", float(tr.flat.abs().max()))
