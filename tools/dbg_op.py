"""Runs the recorded forward descs of a small G one by one (SIMT or tc), with variations of one op (debug aid)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, '.')
from esr_b200 import synth, _capi as capi
from oracle.cem_ops import concat_latent, CEMOracle
from tests.test_gpu_net import build_product_G
dev = torch.device('cuda', 0)
impl, var = sys.argv[1], sys.argv[2]
nb, latent = 1, 'all_layers_HR_downscaled'
wts = synth.make_weights('default', seed=7, nb=nb, latent_input=latent)
lr, z = synth.make_inputs(1, 12, 14, seed=7)
xp = CEMOracle(4).pre_pad(concat_latent(lr, z), 3)
netG = build_product_G(dev, nb, latent, wts, train=True)
G = netG.generated_image_model
B, Cc, h, w = xp.shape
plan = G.plan(B, h, w, 0, False)
plan.run_prep(xp.to(dev).contiguous())
torch.cuda.synchronize()
fn = capi.lib().esr_conv3x3_simt if impl == 'simt' else capi.lib().esr_conv3x3_tc
for i, d in enumerate(plan.descs[:7]):
    if i == 5:
        if var == 'nolat': d.num_kblocks = 6
        if var == 'noepi': d.flags = 0; d.out_bf16 = None
        if var == 'nof32': d.out_f32 = None
        if var == 'nores': d.flags &= ~capi.EPI_RES1
        if var == 'nobf': d.out_bf16 = None
        if var == 'skip': continue
    rc = fn(C.byref(d), capi.stream_ptr())
    torch.cuda.synchronize()
    print(i, 'ok', rc, d.cout_tile, d.cout_tiles, d.pair, d.num_kblocks, hex(d.flags), flush=True)
