"""Fused growth launches vs separate launches per plan shape: forward graph replay time of G+CEM (eval / pre-pad) for a
few (B, h, w); run once with ESR_FUSE_RDB=0 and once with =1 (tools/fuse_shapes.sh)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esr_b200 import synth
from esr_b200.rrdbnet import capture_inference
from tests.test_gpu_net import build_product_G
dev = torch.device('cuda', 0)
wts = synth.make_weights("default", seed=3)
netG = build_product_G(dev, 23, "all_layers_HR_downscaled", wts, train=False)
G = netG.generated_image_model
res = []
for (B, h, w) in [(1, 64, 64), (1, 128, 128), (2, 128, 128), (4, 128, 128), (1, 256, 256), (16, 32, 32), (16, 12, 12)]:
    lr, z = synth.make_inputs(B, h, w, seed=1)
    x = torch.cat([z.contiguous().view(B, 48, h, w), lr], 1).contiguous().to(dev)
    graph, out = capture_inference(G, x, netG._margin_LR, netG._filters)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    hp, wp = h + 20, w + 20
    res.append("%dx%dx%d (%d px, %d tile pairs/layer): %.3f ms" % (B, h, w, B * hp * wp, B * ((hp + 7) // 8) * ((wp + 29) // 30), e0.elapsed_time(e1) / 10))
    del graph
    G._plans.clear()
print("ESR_FUSE_RDB=%s | " % os.environ.get("ESR_FUSE_RDB", "default") + " | ".join(res))
