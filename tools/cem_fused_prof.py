"""Phase breakdown of the single-launch CEM projection (csrc/cem_fused.cu) at BASELINE config 4, from globaltimer stamps
taken by thread 0 of every CTA (ESR_CEM_PROF=1).  Prints, per round, the median / max over CTAs of each phase in us."""
import ctypes as C
import os
import sys

os.environ["ESR_CEM_PROF"] = "1"
os.environ["ESR_CEM_FUSED"] = "1"
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esr_b200 import _capi as capi, cem as pcem  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
f = pcem.CEMnet(pcem.Get_CEM_Config(4))._filters
B, Cc, H, W = 1, 3, 2048, 2048
sets = [(torch.rand(B, Cc, H, W, device=dev), torch.rand(B, Cc, H // 4, W // 4, device=dev), torch.empty(B, Cc, H, W, device=dev)) for _ in range(6)]
ws = torch.empty(2 * B * Cc * (H // 4) * (W // 4), device=dev)
l = capi.lib()
for k in range(13):
    y, x, out = sets[k % 6]
    capi.check(l.esr_cem_project(f, capi.ptr(y), capi.ptr(x), B, Cc, H, W, 0, capi.ptr(out), capi.ptr(ws), capi.stream_ptr()))
torch.cuda.synchronize()
n = 160 * 4 * 8
buf = (C.c_uint64 * n)()
capi.check(l.esr_debug_cem_fused_prof(buf, n))
t = np.array(buf, dtype=np.uint64).reshape(160, 4, 8)[:128, :3].astype(np.float64)
t0 = t[:, 0, 0].min()
names = ["wait TMA", "Down vertical", "Down horizontal + T", "K horizontal + red", "grid.sync", "K vertical + Up vertical", "sync + Up horizontal + store"]
for r in range(3):
    print("round %d: starts at %.2f us (median over CTAs)" % (r, np.median(t[:, r, 0] - t0) / 1e3))
    for i, nm in enumerate(names):
        d = (t[:, r, i + 1] - t[:, r, i]) / 1e3
        print("   %-30s median %6.2f  max %6.2f us" % (nm, np.median(d), d.max()))
print("total (first stamp to last): %.2f us" % ((t[:, 2, 7].max() - t0) / 1e3))
