"""One fused Z-optimisation iteration at BASELINE config 3 for an ncu launch list: set-up + capture, then ITERS graph replays."""
import contextlib, io, os, sys, torch
sys.path.insert(0, '.')
from esr_b200 import synth
from esr_b200.z_optimization import Z_optimizer, SRModelShim
from tests.test_gpu_net import build_product_G
dev = torch.device('cuda', 0)
wts = synth.make_weights("default", seed=3)
netG = build_product_G(dev, 23, "all_layers_HR_downscaled", wts, train=False)
lr, z0 = synth.make_inputs(1, 256, 256, seed=3)
model = SRModelShim(netG)
data = {"LR": lr.to(dev), "Z": torch.zeros_like(z0).to(dev)}
model.feed_data(data)
with torch.no_grad():
    model.fake_H = netG(model.model_input)
with contextlib.redirect_stdout(io.StringIO()):
    zo = Z_optimizer(objective="TV", Z_size=[1024, 1024], model=model, Z_range=1.0, max_iters=int(sys.argv[1]) if len(sys.argv) > 1 else 2, data=data, initial_LR=0.1, batch_size=1)
    zo.optimize()
torch.cuda.synchronize()
print("fused", getattr(zo, "_fused", None) is not None, zo.loss_values)
