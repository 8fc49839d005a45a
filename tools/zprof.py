"""Splits one Z-optimisation iteration (config 3) into forward / backward, GPU time vs host time (debug aid)."""
import sys, time, torch
sys.path.insert(0, '.')
from esr_b200 import synth
from oracle.cem_ops import concat_latent
from tests.test_gpu_net import build_product_G
dev = torch.device('cuda', 0)
wts = synth.make_weights('default', seed=0)
netG = build_product_G(dev, 23, 'all_layers_HR_downscaled', wts)
lr, z = synth.make_inputs(1, 256, 256, seed=3)
lr = lr.to(dev)
zp = z.to(dev).requires_grad_(True)
def it():
    out = netG(concat_latent(lr, zp))
    loss = (out[:, :, :, :-1] - out[:, :, :, 1:]).abs().mean()
    return out, loss
for _ in range(3):
    out, loss = it(); loss.backward(); zp.grad = None
torch.cuda.synchronize()
def timed(fn, n=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return e0.elapsed_time(e1) / n, (t1 - t0) * 1e3 / n, (t2 - t0) * 1e3 / n
print('forward (autograd on): gpu %.2f ms, host-issue %.2f ms, wall %.2f ms' % timed(lambda: it()))
outs = [it() for _ in range(5)]
k = [0]
def bw():
    outs[k[0]][1].backward(); k[0] += 1; zp.grad = None
print('backward:              gpu %.2f ms, host-issue %.2f ms, wall %.2f ms' % timed(bw))
G = netG.generated_image_model
plan = list(G._plans.values())[-1]
print('fwd launches', plan.launches_per_forward(True), 'bwd launches', G.backward_plan(plan).num_launches())
g = torch.cuda.CUDAGraph()
x = concat_latent(lr, zp.detach()).contiguous()
xp = torch.empty(1, 51, 256, 256, device=dev); xp.copy_(x)
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    plan.run_g(xp)
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    plan.run_g(xp)
print('forward G only, CUDA graph: gpu %.2f ms, host %.2f, wall %.2f' % timed(g.replay))
print('forward G only, eager     : gpu %.2f ms, host %.2f, wall %.2f' % timed(lambda: plan.run_g(xp)))
bp = G.backward_plan(plan)
gy = torch.randn(1, 3, 4 * plan.hp, 4 * plan.wp, device=dev)
print('backward G only, eager    : gpu %.2f ms, host %.2f, wall %.2f' % timed(lambda: bp.run(gy)))
g2 = torch.cuda.CUDAGraph()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    bp.run(gy)
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g2):
    bp.run(gy)
print('backward G only, CUDA graph: gpu %.2f ms, host %.2f, wall %.2f' % timed(g2.replay))
