"""conv5-like layer (192->64, RES1+RES2, f32 + bf16 outputs) for profiling (debug aid, GPU box)."""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
from esr_b200 import _capi as capi
from tests.helpers import plain_conv_case, conv_desc, run_conv
dev = torch.device('cuda', 0)
B, H, W = 16, 148, 148
c = plain_conv_case(dev, B, H, W, 192, 64, seed=1, buf_channels=192)
d = conv_desc(c['pc'], B, H, W, c['buf'])
r1 = torch.rand(B, H, W, 64, device=dev); r2 = torch.rand(B, H, W, 64, device=dev)
out32 = torch.zeros(B, H, W, 64, device=dev); outb = torch.zeros(B, H, W, 192, device=dev, dtype=torch.bfloat16)
d.flags = capi.EPI_RES1 | capi.EPI_RES2 | (capi.EPI_F32_BLOCKED if len(sys.argv) > 1 else 0)
d.alpha, d.beta = 0.2, 0.2
d.res1, d.res1_stride, d.res2, d.res2_stride = r1.data_ptr(), 64, r2.data_ptr(), 64
d.out_f32, d.out_f32_stride = out32.data_ptr(), 64
d.out_bf16, d.out_bf16_stride, d.out_bf16_choff = outb.data_ptr(), 192, 0
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run_conv(d, 'tc'); e1.record(); torch.cuda.synchronize()
    print('conv5-like: %.1f us' % (e0.elapsed_time(e1) * 1e3))
