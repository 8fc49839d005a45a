import sys, torch
sys.path.insert(0, '.')
from esr_b200 import synth
from oracle.cem_ops import concat_latent, CEMOracle
from oracle.rrdbnet import RRDBNetOracle
from tests.test_gpu_net import build_product_G
dev = torch.device('cuda', 0)
def rel(a, b): return float((a - b).norm() / (b.norm() + 1e-30))
for nb, kind, latent in ((1, 'default', 'all_layers_HR_downscaled'), (2, 'default', 'all_layers_HR_downscaled'), (3, 'default', 'all_layers_HR_downscaled'), (4, 'default', 'all_layers_HR_downscaled'), (1, 'default', 'first_layer_HR_downscaled'), (2, 'default', 'first_layer_HR_downscaled'), (1, 'default', None)):
    wts = synth.make_weights(kind, seed=7, nb=nb, latent_input=latent)
    lr, z = synth.make_inputs(1, 12, 14, seed=7)
    mi = concat_latent(lr, z) if latent else lr
    xp = CEMOracle(4).pre_pad(mi, 3 if latent else 0)
    with torch.no_grad():
        ref = RRDBNetOracle(wts, nb=nb, latent_input=latent, num_latent_channels=3 if latent else 0).forward(xp)
        emu = RRDBNetOracle(wts, nb=nb, latent_input=latent, num_latent_channels=3 if latent else 0, operand_dtype=torch.bfloat16).forward(xp)
    netG = build_product_G(dev, nb, latent, wts, train=True)
    G = netG.generated_image_model
    for simt in (True,):
        G.debug_simt = simt
        with torch.no_grad():
            y = G(xp.to(dev)).cpu()
        print(nb, kind, latent, 'simt' if simt else 'tc', 'raw G: rel vs fp32', rel(y, ref), 'vs bf16-emu', rel(y, emu), 'emu vs fp32', rel(emu, ref), 'max', float((y - ref).abs().max()), float(ref.abs().max()))
