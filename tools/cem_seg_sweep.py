"""Sweeps the LR-rows-per-warp launch parameter of the x4 streaming CEM kernels (ESR_CEM_SEG_DOWN / ESR_CEM_SEG_UP)
on the config-4 and config-2 shapes; checks that every setting reproduces the default's output.
Run on the GPU box: python tools/cem_seg_sweep.py        """
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esr_b200 import _capi as capi, cem as pcem  # noqa: E402

dev = torch.device("cuda", 0)
f = pcem.CEMnet(pcem.Get_CEM_Config(4))._filters
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def measure(B, C, H, W, crop, y, x, out, ws, n=12):
    def run():
        capi.check(capi.lib().esr_cem_project(f, capi.ptr(y), capi.ptr(x), B, C, H, W, crop, capi.ptr(out), capi.ptr(ws),
                                              capi.stream_ptr()))
    for _ in range(3):
        run()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return 1e3 * ts[len(ts) // 2]


for name, (B, C, H, W, crop) in (("cfg4 1x3x2048^2", (1, 3, 2048, 2048, 0)), ("cfg2 16x3x592^2 crop 40", (16, 3, 592, 592, 40))):
    y = torch.rand(B, C, H, W, device=dev)
    x = torch.rand(B, C, H // 4, W // 4, device=dev)
    out = torch.empty(B, C, H - 2 * crop, W - 2 * crop, device=dev)
    ws = torch.empty(2 * B * C * (H // 4) * (W // 4), device=dev)
    for k in ("ESR_CEM_SEG_DOWN", "ESR_CEM_SEG_UP"):
        os.environ.pop(k, None)
    base = measure(B, C, H, W, crop, y, x, out, ws)
    ref = out.clone()
    print("%s  bps=%s  default %.1f us" % (name, os.environ.get("ESR_CEM_UP_BPS", "4"), base), flush=True)
    for var, segs in (("ESR_CEM_SEG_DOWN", (4, 6, 8, 9, 10, 11, 13, 16, 22, 32)), ("ESR_CEM_SEG_UP", (4, 6, 8, 9, 10, 11, 13, 16))):
        row = []
        for seg in segs:
            os.environ[var] = str(seg)
            out.zero_()
            t = measure(B, C, H, W, crop, y, x, out, ws)
            ok = (out - ref).abs().max().item() <= 1e-6
            row.append("%d:%.1f%s" % (seg, t, "" if ok else "(MISMATCH)"))
        os.environ.pop(var)
        print("  %-17s %s" % (var, "  ".join(row)), flush=True)
