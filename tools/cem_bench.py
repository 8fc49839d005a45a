"""CEM standalone (BASELINE config 4): 2048x2048 x4 SR output, bicubic kernel, vs the HBM roofline."""
import json, os, sys, torch
sys.path.insert(0, '.')
from esr_b200 import _capi as capi, cem as pcem
dev = torch.device('cuda', 0)
net = pcem.CEMnet(pcem.Get_CEM_Config(4))
f = net._filters
B, C, H, W = 1, 3, 2048, 2048
y = torch.rand(B, C, H, W, device=dev); x = torch.rand(B, C, H // 4, W // 4, device=dev)
out = torch.empty_like(y); ws = torch.empty(2 * B * C * (H // 4) * (W // 4), device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
l = capi.lib()
def run():
    capi.check(l.esr_cem_project(f, capi.ptr(y), capi.ptr(x), B, C, H, W, 0, capi.ptr(out), capi.ptr(ws), capi.stream_ptr()))
for _ in range(3): run()
ts = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = sorted(ts)[len(ts) // 2]
byts = 4 * 3 * (2 * H * W + H * W // 16)
peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'] if os.path.exists('MEASURED_PEAKS.json') else 6650.0
print(json.dumps({"cem_2048": {"ms": t, "algorithmic_GBs": byts / t / 1e6, "frac_of_hbm_peak": byts / t / 1e6 / peak, "Mpix_s": H * W / t / 1e3, "l2": "flushed between iterations (256 MiB memset)"}}))
