"""BASELINE config 1 (1x3x64x64 LR + Z, eval/pre-pad, production net): one captured forward, replayed; for an ncu launch list
(SURVEY.md 8d asks for a timeline of its launch overhead; nsys is not installed in this image)."""
import sys, torch
sys.path.insert(0, '.')
from esr_b200 import synth
from esr_b200.rrdbnet import capture_inference
from tests.test_gpu_net import build_product_G
dev = torch.device('cuda', 0)
netG = build_product_G(dev, 23, "all_layers_HR_downscaled", synth.make_weights("default", seed=0))
lr, z = synth.make_inputs(1, 64, 64, seed=1)
x = torch.cat([z.contiguous().view(1, 48, 64, 64), lr], 1).contiguous().to(dev)
graph, out = capture_inference(netG.generated_image_model, x, netG._margin_LR, netG._filters)
for _ in range(3):
    graph.replay()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); graph.replay(); b.record(); torch.cuda.synchronize()
print("config 1 replay: %.1f us" % (a.elapsed_time(b) * 1e3))
