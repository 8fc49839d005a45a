"""One warm Z-optimisation iteration (config 3) between cudaProfilerStart/Stop, for `ncu --profile-from-start off`."""
import sys, torch
sys.path.insert(0, '.')
from esr_b200 import synth
from oracle.cem_ops import concat_latent
from tests.test_gpu_net import build_product_G
dev = torch.device('cuda', 0)
netG = build_product_G(dev, 23, 'all_layers_HR_downscaled', synth.make_weights('default', seed=0))
lr, z = synth.make_inputs(1, 256, 256, seed=3)
lr = lr.to(dev)
zp = z.to(dev).requires_grad_(True)
def it():
    out = netG(concat_latent(lr, zp))
    loss = (out[:, :, :, :-1] - out[:, :, :, 1:]).abs().mean()
    loss.backward(); zp.grad = None
for _ in range(2): it()
torch.cuda.synchronize()
torch.cuda.profiler.start()
it()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
