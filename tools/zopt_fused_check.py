"""Fused single-graph Z-optimisation loop against the generic (autograd + torch.optim.Adam) loop of the same package:
same losses and Z after a few iterations, for every fused objective; then iterations/s of both at BASELINE config 3."""
import contextlib, io, os, sys, time, torch
sys.path.insert(0, '.')
from esr_b200 import synth
from esr_b200.z_optimization import Z_optimizer, SRModelShim
from tests.test_gpu_net import build_product_G
dev = torch.device('cuda', 0)

def run(objective, fused, nb, h, w, iters, bs=1, seed=5):
    os.environ['ESR_ZOPT_FUSED'] = '1' if fused else '0'
    wts = synth.make_weights("default", seed=seed, nb=nb)
    netG = build_product_G(dev, nb, "all_layers_HR_downscaled", wts, train=False)
    lr, z0 = synth.make_inputs(1, h, w, seed=seed)
    model = SRModelShim(netG)
    data = {"LR": lr.repeat(bs, 1, 1, 1).to(dev), "Z": (0.5 * z0).repeat(bs, 1, 1, 1).to(dev)}
    if 'increase' in objective or 'decrease' in objective:
        data['STD_increment'] = 0.02
    model.feed_data(data)
    with torch.no_grad():
        model.fake_H = netG(model.model_input)
    with contextlib.redirect_stdout(io.StringIO()):
        zo = Z_optimizer(objective=objective, Z_size=[4 * h, 4 * w], model=model, Z_range=1.0, max_iters=iters, data=data, initial_LR=0.1, batch_size=bs)
        if bs > 1:
            zo.random_Z_inits = False
            zo.Z_model.Z.data.copy_(0.3 * torch.randn(zo.Z_model.Z.shape, generator=torch.Generator().manual_seed(1)).to(dev))
        Z = zo.optimize()
        used = getattr(zo, '_fused', None) is not None
        torch.cuda.synchronize()
        t0 = time.perf_counter(); Z2 = zo.optimize(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return zo.loss_values, zo.latest_Z_loss_values, Z2.cpu(), used, iters / dt, model.fake_H.cpu()

for obj, bs in (("TV", 1), ("max_STD", 1), ("min_STD", 1), ("STD_increase", 1), ("STD_decrease", 1), ("TV", 2)):
    a = run(obj, True, 2, 12, 16, 4, bs)
    b = run(obj, False, 2, 12, 16, 4, bs)
    rel = max(abs(x - y) / max(abs(y), 1e-12) for x, y in zip(a[0], b[0]))
    print("%-13s bs %d fused path used %s | loss rel diff %.2e | latest %s vs %s | mean |dZ| %.2e | fake_H max diff %.2e" % (
        obj, bs, a[3], rel, ["%.5f" % v for v in a[1]], ["%.5f" % v for v in b[1]], float((a[2] - b[2]).abs().mean()), float((a[5] - b[5]).abs().max())))
for fused in (True, False):
    r = run("TV", fused, 23, 256, 256, 10, 1, seed=3)
    print("config 3: fused %s -> %.1f it/s, losses %.5f -> %.5f" % (fused, r[4], r[0][0], r[0][-1]))
