"""Fixed vs per-iteration cost of Z_optimizer.optimize at BASELINE config 3, fused and generic loops."""
import contextlib, io, os, sys, time, torch
sys.path.insert(0, '.')
from esr_b200 import synth, z_optimization as zmod
from esr_b200.z_optimization import Z_optimizer, SRModelShim
from tests.test_gpu_net import build_product_G
dev = torch.device('cuda', 0)
wts = synth.make_weights("default", seed=3)
netG = build_product_G(dev, 23, "all_layers_HR_downscaled", wts, train=False)
lr, z0 = synth.make_inputs(1, 256, 256, seed=3)
n_built = [0]
orig = zmod._FusedZLoop.__init__
def counting(self, *a, **k):
    n_built[0] += 1
    return orig(self, *a, **k)
zmod._FusedZLoop.__init__ = counting
for fused in ("1", "0"):
    os.environ['ESR_ZOPT_FUSED'] = fused
    model = SRModelShim(netG)
    data = {"LR": lr.to(dev), "Z": torch.zeros_like(z0).to(dev)}
    model.feed_data(data)
    with torch.no_grad():
        model.fake_H = netG(model.model_input)
    with contextlib.redirect_stdout(io.StringIO()):
        zo = Z_optimizer(objective="TV", Z_size=[1024, 1024], model=model, Z_range=1.0, max_iters=2, data=data, initial_LR=0.1, batch_size=1)
        zo.optimize()
        res = []
        for iters in (10, 50, 10, 50):
            zo.max_iters = iters
            torch.cuda.synchronize(); t0 = time.perf_counter(); zo.optimize(); torch.cuda.synchronize()
            res.append((iters, (time.perf_counter() - t0) * 1e3))
    print("fused", fused, "graphs built", n_built[0], ["%d it: %.1f ms" % r for r in res],
          "per-iteration %.3f ms" % ((res[1][1] - res[0][1]) / 40), "fixed %.2f ms" % (res[0][1] - 10 * (res[1][1] - res[0][1]) / 40))
    if fused == "1":
        f = zo._fused
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        evs[0].record()
        for _ in range(20): f.graph.replay()
        evs[1].record(); torch.cuda.synchronize()
        print("graph replay alone: %.3f ms / iteration" % (evs[0].elapsed_time(evs[1]) / 20))
