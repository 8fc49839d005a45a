"""Per-role cycle breakdown of single conv launches (debug aid, GPU box)."""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
from esr_b200 import _capi as capi
from tests.helpers import plain_conv_case, conv_desc, run_conv
dev = torch.device('cuda', 0)
l = capi.lib()
l.esr_debug_set_profile_buffer.argtypes = [C.c_void_p]
prof = torch.zeros(148, 16, dtype=torch.int64, device=dev)
for (cin, cout, lrelu) in ((64, 32, True), (160, 32, True), (192, 64, False), (192, 64, None)):
    B, H, W = 16, 148, 148
    c = plain_conv_case(dev, B, H, W, cin, cout, seed=1, buf_channels=192)
    buf = c['buf']
    d = conv_desc(c['pc'], B, H, W, buf)
    out32 = torch.zeros(B, H, W, 64, device=dev) if not lrelu else None
    if lrelu is None:
        r1 = torch.rand(B, H, W, 64, device=dev); r2 = torch.rand(B, H, W, 64, device=dev)
        outb = torch.zeros(B, H, W, 192, device=dev, dtype=torch.bfloat16)
        d.flags = capi.EPI_RES1 | capi.EPI_RES2 | capi.EPI_F32_BLOCKED
        d.alpha, d.beta = 0.2, 0.2
        d.res1, d.res1_stride, d.res2, d.res2_stride = r1.data_ptr(), 64, r2.data_ptr(), 64
        d.out_f32, d.out_f32_stride = out32.data_ptr(), 64
        d.out_bf16, d.out_bf16_stride, d.out_bf16_choff = outb.data_ptr(), 192, 0
    elif lrelu:
        d.flags = capi.EPI_LRELU
        d.out_bf16, d.out_bf16_stride, d.out_bf16_choff = buf.data_ptr(), 192, 160
    else:
        d.out_f32, d.out_f32_stride = out32.data_ptr(), 64
    l.esr_debug_set_profile_buffer(C.c_void_p(prof.data_ptr()))
    for it in range(3):
        prof.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run_conv(d, 'tc'); e1.record(); torch.cuda.synchronize()
    p = prof.double().mean(0).cpu()
    print('cin %d cout %d: %.1f us' % (cin, cout, e0.elapsed_time(e1) * 1e3))
    print('  producer: total %.0f wait_empty %.0f n_kb %.0f -> per kb total %.0f wait %.0f' % (p[0], p[1], p[2], p[0] / p[2], p[1] / p[2]))
    print('  mma: total %.0f wait_acc_empty %.0f wait_full %.0f wait_w %.0f' % (p[3], p[4], p[5], p[6]))
    print('  epi(warp2): total %.0f wait_acc_full %.0f ld %.0f shfl %.0f tiles %.0f -> per tile total %.0f wait %.0f ld %.0f shfl %.0f store %.0f' % (
        p[7], p[8], p[9], p[10], p[11], p[7] / p[11], p[8] / p[11], p[9] / p[11], p[10] / p[11], (p[7] - p[8] - p[9] - p[10]) / p[11]))
    q = prof.cpu().double()
    t0 = q[:, 12].min()
    rel = lambda c: ((q[:, c] - t0) / 1e3)
    print('  timeline us (min/mean/max over CTAs): entry %.1f/%.1f/%.1f  setup_done %.1f/%.1f/%.1f  mma_end %.1f/%.1f/%.1f  epi_end %.1f/%.1f/%.1f' % (
        rel(12).min(), rel(12).mean(), rel(12).max(), rel(13).min(), rel(13).mean(), rel(13).max(),
        rel(14).min(), rel(14).mean(), rel(14).max(), rel(15).min(), rel(15).mean(), rel(15).max()))
