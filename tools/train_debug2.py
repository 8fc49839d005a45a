import os, sys, torch
sys.path.insert(0, '.')
os.environ.setdefault('CUDA_LAUNCH_BLOCKING', '1')
from esr_b200 import synth
from oracle.cem_ops import concat_latent
from tests.test_gpu_net import build_product_G
dev = torch.device('cuda', 0)
nb, B, h, train = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4] == 'train'
wts = synth.make_weights("kaiming", seed=0, nb=nb)
netG = build_product_G(dev, nb, "all_layers_HR_downscaled", wts, train=train)
lr, z = synth.make_inputs(B, h, h, seed=0)
x = concat_latent(lr, z).to(dev).requires_grad_(True)
out = netG(x); torch.cuda.synchronize(); print("forward ok")
out.sum().backward(); torch.cuda.synchronize(); print("autograd backward ok", nb, B, h, train, float(x.grad.abs().max()))
