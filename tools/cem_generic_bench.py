"""Times the generic separable CEM kernels (any scale / tap count): x2 and x3 bicubic, and x4 with a mild blur, at
16 x 3 planes of 148 x 148 LR cells.  Run on the GPU box: python tools/cem_generic_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esr_b200 import _capi as capi, cem as pcem  # noqa: E402

dev = torch.device("cuda", 0)
for sf, kern in ((2, None), (3, None), (4, "blurry_cubic_1"), (4, None)):
    net = pcem.CEMnet(pcem.Get_CEM_Config(sf), upscale_kernel=kern)
    f = net._filters
    B, C, h, w = 16, 3, 148, 148
    H, W = sf * h, sf * w
    crop = sf * int(net.invalidity_margins_LR)
    y = torch.rand(B, C, H, W, device=dev)
    x = torch.rand(B, C, h, w, device=dev)
    out = torch.empty(B, C, H - 2 * crop, W - 2 * crop, device=dev)
    ws = torch.empty(2 * B * C * h * w, device=dev)

    def run():
        capi.cem_call("project", f, capi.ptr(y), capi.ptr(x), B, C, H, W, crop, capi.ptr(out), capi.ptr(ws), capi.stream_ptr())
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    print("x%d %-16s ds %2d inv %2d crop %3d: project %.3f ms" % (sf, kern or "bicubic", len(net._ds_1d), len(net._inv_1d), crop,
                                                                 e0.elapsed_time(e1) / 20), flush=True)
