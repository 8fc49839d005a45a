"""Per-role cycle breakdown of single pair-kernel conv launches (debug aid, GPU box; build with -DESR_PROFILE_ROLES)."""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
from esr_b200 import _capi as capi
from tests.helpers import plain_conv_case, conv_desc, run_conv
dev = torch.device('cuda', 0)
l = capi.lib()
l.esr_debug_set_profile_buffer.argtypes = [C.c_void_p]
prof = torch.zeros(148, 16, dtype=torch.int64, device=dev)
B, H, W = 16, 148, 148
for (cin, cout, ct, kind) in ((64, 32, 32, 'trunk'), (160, 32, 32, 'trunk'), (192, 64, 64, 'res'), (64, 64, 64, 'hr')):
    if kind == 'hr':
        B, H, W = 4, 592, 592
    c = plain_conv_case(dev, B, H, W, cin, cout, seed=1, buf_channels=192 if kind != 'hr' else 64, pair=True, cout_tile=ct)
    buf = c['buf']
    d = conv_desc(c['pc'], B, H, W, buf)
    keep = []
    if kind == 'res':
        r1 = torch.rand(B, H, W, 64, device=dev); out32 = torch.zeros(B, H, W, 64, device=dev)
        outb = torch.zeros(B, H, W, 192, device=dev, dtype=torch.bfloat16)
        d.flags = capi.EPI_RES1 | capi.EPI_F32_BLOCKED
        d.alpha = 0.2
        d.res1, d.res1_stride = r1.data_ptr(), 64
        d.out_f32, d.out_f32_stride = out32.data_ptr(), 64
        d.out_bf16, d.out_bf16_stride, d.out_bf16_choff = outb.data_ptr(), 192, 0
        keep = [r1, out32, outb]
    elif kind == 'trunk':
        d.flags = capi.EPI_LRELU
        d.out_bf16, d.out_bf16_stride, d.out_bf16_choff = buf.data_ptr(), 192, 160
    else:
        outb = torch.zeros(B, H, W, 64, device=dev, dtype=torch.bfloat16)
        d.flags = capi.EPI_LRELU | capi.EPI_OUT_F16
        d.out_bf16, d.out_bf16_stride, d.out_bf16_choff = outb.data_ptr(), 64, 0
        keep = [outb]
    l.esr_debug_set_profile_buffer(C.c_void_p(prof.data_ptr()))
    for it in range(3):
        prof.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run_conv(d, 'tc'); e1.record(); torch.cuda.synchronize()
    q = prof.cpu().double()
    lead, peer = q[0::2], q[1::2]
    print('%s cin %d cout %d (%dx%dx%d): %.1f us event-timed' % (kind, cin, cout, B, H, W, e0.elapsed_time(e1) * 1e3))
    for nm, p in (('leader', lead.mean(0)), ('peer', peer.mean(0))):
        print('  %s producer: total %.0f wait_empty %.0f n_kb %.0f -> per kb total %.0f wait %.0f' % (nm, p[0], p[1], p[2], p[0] / max(p[2], 1), p[1] / max(p[2], 1)))
    p = lead.mean(0)
    print('  mma: total %.0f wait_acc_empty %.0f wait_full %.0f wait_w %.0f -> busy %.0f' % (p[3], p[4], p[5], p[9], p[3] - p[4] - p[5] - p[9]))
    p = q.mean(0)
    print('  epi(warp2): total %.0f wait_acc_full %.0f tiles %.0f -> per tile total %.0f wait %.0f work %.0f' % (
        p[7], p[8], p[11], p[7] / p[11], p[8] / p[11], (p[7] - p[8]) / p[11]))
    t0 = q[:, 12].min()
    rel = lambda col, sel=q: ((sel[:, col] - t0) / 1e3)
    f = lambda v: '%.1f/%.1f/%.1f' % (v.min(), v.mean(), v.max())
    print('  timeline us (min/mean/max): entry %s setup_done %s tma_start %s first_full(leader) %s mma_end(leader) %s epi_end %s' % (
        f(rel(12)), f(rel(13)), f(rel(6)), f(rel(10, lead)), f(rel(14, lead)), f(rel(15))))
