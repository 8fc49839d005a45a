"""Breakdown of the generator training step at BASELINE config 5 (16 x 3x32x32 LR, nb=23, train mode): wall-clock with a
device sync after every phase (host + device), then the device-only time of the weight-gradient launches."""
import sys, time, torch
sys.path.insert(0, '.')
from esr_b200 import synth
from esr_b200.training import GeneratorTrainer
from oracle.cem_ops import concat_latent
from tests.test_gpu_net import build_product_G
dev = torch.device('cuda', 0)
wts = synth.make_weights("kaiming", seed=0)
netG = build_product_G(dev, 23, "all_layers_HR_downscaled", wts, train=True)
G = netG.generated_image_model
for p in G.parameters(): p.requires_grad_(True)
tr = GeneratorTrainer(netG)
opt = torch.optim.Adam(G.parameters(), lr=1e-4)
lr, z = synth.make_inputs(16, 32, 32, seed=0)
mi = concat_latent(lr, z).to(dev)
target = torch.rand(16, 3, 128, 128).to(dev)
def t():
    torch.cuda.synchronize(); return time.perf_counter()
acc = {}
for it in range(6):
    t0 = t(); fake = tr.forward(mi); t1 = t()
    loss = (fake - target).abs().mean(); loss.backward(); t2 = t()
    tr.backward(fake.grad); t3 = t()
    opt.step(); t4 = t()
    if it >= 2:
        for k, v in (("forward (incl. re-pack after the update)", t1 - t0), ("loss", t2 - t1), ("dgrad + wgrad", t3 - t2), ("adam", t4 - t3)):
            acc[k] = acc.get(k, 0) + v / 4
print({k: "%.2f ms" % (v * 1e3) for k, v in acc.items()}, "total %.2f ms" % (1e3 * sum(acc.values())))
# device time of the pieces
plan, filters, margin = None, None, None
fake = tr.forward(mi); plan, filters, margin = tr._state
from esr_b200.backward import generator_backward_eager
bp = tr._backward_plan(plan)
g = torch.randn_like(fake)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
torch.cuda.synchronize()
ev[0].record(); generator_backward_eager(plan, bp, filters, margin, g); ev[1].record()
tr._state = (plan, filters, margin)
torch.cuda.synchronize()
ev[2].record(); tr.backward(g); ev[3].record(); torch.cuda.synchronize()
print("device: dgrad %.2f ms, dgrad + wgrad %.2f ms" % (ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); G.engine().pack(G._packed_params); G._dgrad.pack(G._packed_params); e1.record(); torch.cuda.synchronize()
t0 = time.perf_counter(); G.engine().pack(G._packed_params); G._dgrad.pack(G._packed_params); t1 = time.perf_counter(); torch.cuda.synchronize()
print("re-pack: device %.2f ms, host enqueue %.2f ms" % (e0.elapsed_time(e1), (t1 - t0) * 1e3))
