"""Bisects the x4 CEM streaming kernels on a small case: ESR_CEM_DEBUG=1 (new Down + old K+Up), =2 (old Down + new K+Up), unset
(both new); prints max error against the old kernels' result and the ring-wait timeout record."""
import ctypes as C, os, sys, torch
sys.path.insert(0, '.')
from esr_b200 import _capi as capi, cem as pcem
dev = torch.device('cuda', 0)
f = pcem.CEMnet(pcem.Get_CEM_Config(4))._filters
l = capi.lib()
shapes = [(2, 3, 48, 64, 0), (1, 3, 2048, 2048, 0), (2, 3, 592, 592, 40), (1, 3, 112, 144, 40), (1, 3, 64, 1024, 0), (16, 3, 592, 592, 40), (1, 3, 520, 48, 8)]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])] if int(sys.argv[1]) > 0 else [shapes[-int(sys.argv[1])]]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
from oracle.cem_ops import CEMOracle
ora = CEMOracle(4)
for (B, Cc, H, W, crop) in shapes:
    g = torch.Generator().manual_seed(1)
    y = torch.rand(B, Cc, H, W, generator=g); x = torch.rand(B, Cc, H // 4, W // 4, generator=g)
    yd, xd = y.to(dev), x.to(dev)
    out = torch.full((B, Cc, H - 2 * crop, W - 2 * crop), float('nan'), device=dev)
    ws = torch.empty(2 * B * Cc * (H // 4) * (W // 4), device=dev)
    for _ in range(reps):
        capi.check(l.esr_cem_project(f, capi.ptr(yd), capi.ptr(xd), B, Cc, H, W, crop, capi.ptr(out), capi.ptr(ws), capi.stream_ptr()))
    torch.cuda.synchronize()
    rec = (C.c_uint32 * 4)()
    capi.check(l.esr_debug_cem_timeout(rec))
    ref = ora.project(y, x)
    if crop:
        ref = ref[:, :, crop:-crop, crop:-crop]
    err = (out.cpu() - ref).abs()
    e2 = err.nan_to_num(9e9)
    loc = [int(v) for v in torch.unravel_index(e2.argmax(), e2.shape)]
    bad = (e2 > 1e-4)
    rows = sorted(set((bad.nonzero()[:, 2] // 4).tolist()))[:12]
    cols = sorted(set((bad.nonzero()[:, 3] // 4).tolist()))[:12]
    if int(bad.sum()):
        nz = bad.nonzero()
        print("   bad planes", sorted(set((nz[:, 0] * Cc + nz[:, 1]).tolist())), "HR rows", sorted(set(nz[:, 2].tolist())), "HR cols", sorted(set(nz[:, 3].tolist())))
    print((B, Cc, H, W, crop), "max err vs oracle", float(e2.max()), "at", loc, "n bad", int(bad.sum()), "bad LR rows", rows, "bad LR cols", cols,
          "nan", int(out.isnan().sum()), "timeout", list(rec))
