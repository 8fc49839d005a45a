"""Where one Z-optimisation iteration (BASELINE config 3) spends its time: torch.profiler over the reference-shaped
loop of z_optimization.Z_optimizer.  Prints GPU busy time per iteration, the top kernels and the wall time."""
import contextlib
import io
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esr_b200 import synth  # noqa: E402
from esr_b200.z_optimization import Z_optimizer, SRModelShim  # noqa: E402
from tests.test_gpu_net import build_product_G  # noqa: E402

dev = torch.device("cuda", 0)
wts = synth.make_weights("default", seed=0)
netG = build_product_G(dev, 23, "all_layers_HR_downscaled", wts)
lr3, z3 = synth.make_inputs(1, 256, 256, seed=3)
model = SRModelShim(netG)
data = {"LR": lr3.to(dev), "Z": torch.zeros_like(z3).to(dev)}
model.feed_data(data)
with torch.no_grad():
    model.fake_H = netG(model.model_input)
with contextlib.redirect_stdout(io.StringIO()):
    zo = Z_optimizer(objective="TV", Z_size=[1024, 1024], model=model, Z_range=1.0, max_iters=3, data=data, initial_LR=0.1, batch_size=1)
    zo.optimize()
    torch.cuda.synchronize()
    n = 10
    zo.max_iters = n
    t0 = time.perf_counter()
    zo.optimize()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        zo.optimize()
        torch.cuda.synchronize()
print("wall per iteration (unprofiled): %.2f ms" % wall)
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
busy = sum(e.device_time for e in ev) / n / 1e3
print("GPU busy per iteration: %.2f ms over %d device events / iteration" % (busy, len(ev) // n))
agg = {}
for e in ev:
    k = e.name[:70]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += e.device_time
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print("%8.3f ms/it  %5d/it  %s" % (t / n / 1e3, c // n, k))
