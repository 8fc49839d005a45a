"""Debug aid (GPU box): compares backward intermediates with oracle autograd."""
import sys, torch, torch.nn.functional as F
sys.path.insert(0, '.')
from esr_b200 import synth
from oracle.cem_ops import concat_latent
from oracle.rrdbnet import GCEMOracle
from tests.test_gpu_net import build_product_G
dev = torch.device('cuda', 0)
nb, latent, train, kind = 1, "all_layers_HR_downscaled", False, "default"
impl = sys.argv[1] if len(sys.argv) > 1 else 'simt'
wts = synth.make_weights(kind, seed=7, nb=nb, latent_input=latent)
lr, z = synth.make_inputs(1, 12, 14, seed=7)
gout = torch.randn(1, 3, 48, 56, generator=torch.Generator().manual_seed(3))
ora = GCEMOracle(wts, pre_pad=True, nb=nb, latent_input=latent)
net, cem = ora.net, ora.cem
mi = concat_latent(lr, z)
xp = cem.pre_pad(mi, 3)
zpk, x = xp[:, :-3], xp[:, -3:]
b, _, h, w = x.shape
lat_hr = zpk.reshape(b, -1, 4 * h, 4 * w).clone().requires_grad_(True)
lat_lr = F.interpolate(lat_hr.detach(), scale_factor=0.25, mode='bilinear', align_corners=False).clone().requires_grad_(True)
f = net.conv(torch.cat([lat_lr, x], 1), "model.0", act=False); f.retain_grad()
t = net.trunk(torch.cat([lat_lr, f], 1), lat_lr); t.retain_grad()
u = t
ups = []
for k in range(2):
    u = F.interpolate(u, scale_factor=2, mode='nearest'); u = net.conv(u, "model.%d.1" % (2 + k), act=True); u.retain_grad(); ups.append(u)
v = net.conv(torch.cat([lat_hr, u], 1), "model.4", act=True); v.retain_grad()
y = net.conv(torch.cat([lat_hr, v], 1), "model.6", act=False); y.retain_grad()
out = cem.unpad_HR(cem.project(y, x))
(out * gout).sum().backward()

netG = build_product_G(dev, nb, latent, wts, train=False)
G = netG.generated_image_model
G.debug_simt = impl == 'simt'
zp = z.clone().to(dev).requires_grad_(True)
o = netG(concat_latent(lr.to(dev), zp))
plan = list(G._plans.values())[-1]
def nchw(t): return t.float().permute(0, 3, 1, 2).cpu()
def rel(a, b): return float((a - b).norm() / (b.norm() + 1e-30))
torch.cuda.synchronize()
V2a = nchw(plan.V2); V1a = nchw(plan.V1)
print('after fwd: y', rel(plan.y.cpu(), y.detach()), 'V2', rel(V2a[:, :64] + V2a[:, 64:], v.detach()), 'V1', rel(V1a[:, :64] + V1a[:, 64:], ups[1].detach()),
      'U1', rel(nchw(plan.U[1])[:, :64, ::2, ::2] + nchw(plan.U[1])[:, 64:, ::2, ::2], ups[0].detach()), 'nplans', len(G._plans), [k for k in G._plans])
(o * gout.to(dev)).sum().backward()
torch.cuda.synchronize()
V2b = nchw(plan.V2)
print('V2 changed by backward:', float((V2a - V2b).abs().max()))
bp = list(G._bplans.values())[-1][1]
def rel(a, b): return float((a - b).norm() / (b.norm() + 1e-30))
def blocked(t, C, H, W): return t.view(1, C // 8, H, W, 8).permute(0, 1, 4, 2, 3).reshape(1, C, H, W).cpu()
H4, W4 = 4 * h, 4 * w
print('out fwd', rel(o.detach().cpu(), out.detach()))
print('g_y', rel(bp.g_y.cpu(), y.grad))
GH = blocked(bp.GH, 96, H4, W4)
print('GH main (d v, fp32 unmasked)', rel(GH[:, :64], v.grad), 'lat part', float(GH[:, 64:73].abs().mean()))
print('   per-channel rel', [round(rel(GH[:, c], v.grad[:, c]), 3) for c in (0, 1, 31, 32, 63)])
print('   interior rel', rel(GH[:, :64, 4:-4, 4:-4], v.grad[:, :, 4:-4, 4:-4]))
e6 = bp.E6.float().permute(0, 3, 1, 2).cpu()
print('E6 hi+lo vs g_y', rel(e6[:, 0:3] + e6[:, 3:6], y.grad), 'hi2', rel(e6[:, 6:9], e6[:, 0:3]), 'pad', float(e6[:, 9:].abs().max()))
dv_ct = F.conv_transpose2d(y.grad, wts['model.6.weight'], padding=1)[:, 3:]
print('v.grad vs conv_transpose', rel(dv_ct, v.grad))
V2 = plan.V2.float().permute(0, 3, 1, 2).cpu()
print('V2 vs v', rel(V2[:, :64] + V2[:, 64:], v.detach()), 'sign agreement', float(((V2[:, :64] > 0) == (v > 0)).float().mean()))
gv = bp.GV.float().permute(0, 3, 1, 2).cpu(); gv = gv[:, :64] + gv[:, 64:]
dv_pre = v.grad * (v > 0).float() + 0.2 * v.grad * (v <= 0).float()
print('d(pre-act HR_conv0 out) [GV]', rel(gv, dv_pre))
gv1 = bp.GV1.float().permute(0, 3, 1, 2).cpu(); gv1 = gv1[:, :64] + gv1[:, 64:]
du_pre = ups[1].grad * torch.where(ups[1] > 0, 1.0, 0.2)
print('d(pre-act upconv2 out) [GV1]', rel(gv1, du_pre))
gv0 = bp.GVu[0].float().permute(0, 3, 1, 2).cpu(); gv0 = gv0[:, :64] + gv0[:, 64:]
du0_pre = ups[0].grad * torch.where(ups[0] > 0, 1.0, 0.2)
print('d(pre-act upconv1 out) [GV0]', rel(gv0, du0_pre))
gsc = blocked(bp.Gsc, 64, h, w)
print('G_sc', rel(gsc, t.grad))
print('g_z_hr', rel(bp.g_z_hr.cpu(), lat_hr.grad))
print('g_z_lr', rel(bp.g_z_lr.cpu(), lat_lr.grad))
GF = bp.GF32.view(1, 4 * 224 // 8, h, w, 8).permute(0, 1, 4, 2, 3).reshape(1, 896, h, w).cpu()
print('d fea (frame0 x0 + Gsc)', rel(GF[:, :64] + gsc, f.grad))
print('total dz', rel(zp.grad.cpu(), None if False else (lambda: 0)() or zp.grad.cpu()))
