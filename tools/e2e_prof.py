"""Where the e2e step loses time against the device-resident step (debug aid, GPU box)."""
import sys, torch
sys.path.insert(0, '.')
from esr_b200 import synth, networks, cem as pcem
from esr_b200.parallel import HostPipeline
dev = torch.device('cuda', 0)
opt = {"gpu_ids": None, "is_train": False, "datasets": {"train": {"patch_size": 256}},
       "network_G": dict(which_model_G="RRDB_net", CEM_arch=1, latent_input="all_layers", latent_input_domain="HR_downscaled",
                         latent_channels=3, norm_type=None, mode="CNA", nf=64, nb=23, in_nc=3, out_nc=3, gc=32, scale=4)}
netG = networks.define_G(opt, CEM=pcem.CEMnet(pcem.Get_CEM_Config(4)), num_latent_channels=3).to(dev).eval()
for p in netG.parameters(): p.requires_grad_(False)
lr, z = synth.make_inputs(16, 128, 128, seed=0)
host_in = torch.cat([z.contiguous().view(16, 48, 128, 128), lr], 1).contiguous().pin_memory()
host_out = torch.empty(16, 3, 512, 512).pin_memory()
x = host_in.to(dev)
def timed(fn, n=10, w=5):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
with torch.no_grad():
    print('compute only (eager netG on device tensor): %.2f ms' % timed(lambda: netG(x)))
    xs = torch.empty_like(x)
    def h2d_sync(): xs.copy_(host_in, non_blocking=True); netG(xs)
    print('same-stream H2D + compute: %.2f ms' % timed(h2d_sync))
    def d2h_sync(): host_out.copy_(netG(x), non_blocking=True)
    print('compute + same-stream D2H: %.2f ms' % timed(d2h_sync))
    print('H2D alone: %.2f ms   D2H alone: %.2f ms' % (timed(lambda: xs.copy_(host_in, non_blocking=True)), timed(lambda: host_out.copy_(x[:, :3].repeat(1, 1, 4, 4)[:16], non_blocking=True))))
g0, out0 = netG.capture(x.clone(), slot=0)
print('captured graph replay only: %.2f ms' % timed(g0.replay))
g1, out1 = netG.capture(x.clone(), slot=1)
def alt():
    g0.replay(); g1.replay()
print('two captured graphs alternating (2 buffer sets): %.2f ms per replay' % (timed(alt) / 2))
s_out = torch.cuda.Stream()
def replay_d2h():
    g0.replay()
    ev = torch.cuda.Event(); ev.record()
    with torch.cuda.stream(s_out):
        s_out.wait_event(ev)
        host_out.copy_(out0, non_blocking=True)
print('graph replay + D2H on side stream: %.2f ms' % timed(replay_d2h))
torch.cuda.synchronize()
s_in = torch.cuda.Stream()
xs2 = [x.clone(), x.clone()]
def h2d_replay():
    with torch.cuda.stream(s_in):
        xs2[0].copy_(host_in, non_blocking=True)
    g0.replay()
print('graph replay + concurrent H2D (other buffer): %.2f ms' % timed(h2d_replay))
torch.cuda.synchronize()
for g in (False, True):
    pipe = HostPipeline(netG, chunk=16, use_graph=g)
    def f(): pipe(host_in, host_out)
    t = timed(f)
    pipe.wait()
    print('HostPipeline graph=%s: %.2f ms' % (g, t))
