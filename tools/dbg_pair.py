"""Forward of a small G through the pair kernels with per-op synchronisation (debug aid, GPU box)."""
import os, sys, torch
os.environ.setdefault("ESR_SEQ_SYNC", "1")
sys.path.insert(0, '.')
from esr_b200 import synth
from oracle.cem_ops import concat_latent, CEMOracle
from oracle.rrdbnet import RRDBNetOracle
from tests.test_gpu_net import build_product_G
dev = torch.device('cuda', 0)
def rel(a, b): return float((a - b).norm() / (b.norm() + 1e-30))
nb, latent = 1, 'all_layers_HR_downscaled'
wts = synth.make_weights('default', seed=7, nb=nb, latent_input=latent)
lr, z = synth.make_inputs(1, 12, 14, seed=7)
mi = concat_latent(lr, z)
xp = CEMOracle(4).pre_pad(mi, 3)
with torch.no_grad():
    ref = RRDBNetOracle(wts, nb=nb, latent_input=latent, num_latent_channels=3).forward(xp)
netG = build_product_G(dev, nb, latent, wts, train=True)
G = netG.generated_image_model
for simt in (True, False):
    G.debug_simt = simt
    with torch.no_grad():
        y = G(xp.to(dev)).cpu()
    print('simt' if simt else 'tc', 'rel vs fp32', rel(y, ref), 'max', float((y - ref).abs().max()))
