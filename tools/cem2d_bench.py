"""Times the general 2-D CEM stencils (csrc/cem2d.cu) next to the separable kernels on the BASELINE config-2 CEM
shape (16 x 3 x 592 x 592 HR, crop 40).  Run on the GPU box: python tools/cem2d_bench.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esr_b200 import _capi as capi, cem as pcem  # noqa: E402


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    dev = torch.device("cuda", 0)
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "cem_nondefault.npz"))
    B, C, H, W, crop = 16, 3, 592, 592, 40
    y = torch.rand(B, C, H, W, device=dev)
    x = torch.rand(B, C, H // 4, W // 4, device=dev)
    gout = torch.randn(B, C, H - 2 * crop, W - 2 * crop, device=dev)
    default = pcem.CEMnet(pcem.Get_CEM_Config(4))
    cases = [("bicubic (separable, streaming x4)", default._filters),
             ("bicubic through the 2-D stencils", capi.CemFilterBank2D(4, default.pre_stride, default.ds_kernel, default.inv_hTh))]
    for name in ("blur1_x4", "blur2_x4", "aniso13_x4", "aniso21s_x4"):
        k = g[name + "_kernel"]
        net = pcem.CEMnet(pcem.Get_CEM_Config(4), upscale_kernel=str(k) if k.dtype.kind in "US" else k)
        cases.append(("%s ds %d inv %d %s" % (name, net.ds_kernel.shape[0], net.inv_hTh.shape[0],
                                               "separable" if net.separable else "2-D"), net._filters))
    out = torch.empty(B, C, H - 2 * crop, W - 2 * crop, device=dev)
    ws = torch.empty(2 * B * C * (H // 4) * (W // 4), device=dev)
    gy = torch.empty(B, C, H, W, device=dev)
    ws2 = torch.empty(B * C * (H * W + H * (W // 4) + 2 * (H // 4) * (W // 4)), device=dev)
    for name, f in cases:
        fwd = timed(lambda: capi.cem_call("project", f, capi.ptr(y), capi.ptr(x), B, C, H, W, crop, capi.ptr(out), capi.ptr(ws),
                                          capi.stream_ptr()))
        bwd = timed(lambda: capi.cem_call("project_bwd", f, capi.ptr(gout), B, C, H, W, crop, capi.ptr(gy), capi.ptr(ws2),
                                          capi.stream_ptr()), n=5)
        print("%-48s project %8.3f ms   project_bwd %8.3f ms" % (name, fwd, bwd), flush=True)


if __name__ == "__main__":
    main()
