"""TMA tile-fill rate: the same 32-channel conv reading a dense 32-channel buffer vs a 32-channel window of a
192-channel NHWC buffer (debug aid, GPU box; library built with -DESR_PROFILE_ROLES)."""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
from esr_b200 import _capi as capi
from tests.helpers import plain_conv_case, conv_desc, run_conv
dev = torch.device('cuda', 0)
l = capi.lib()
l.esr_debug_set_profile_buffer.argtypes = [C.c_void_p]
prof = torch.zeros(148, 16, dtype=torch.int64, device=dev)
B, H, W = 16, 148, 148
for bc in (32, 64, 192):
    c = plain_conv_case(dev, B, H, W, 32, 32, seed=1, buf_channels=bc, pair=True, cout_tile=32)
    # repeat the same K block 6 times so the launch is long enough to be in steady state
    pc = c['pc']
    d = conv_desc(pc, B, H, W, c['buf'])
    outb = torch.zeros(B, H, W, 32, device=dev, dtype=torch.bfloat16)
    d.flags = capi.EPI_LRELU
    d.out_bf16, d.out_bf16_stride, d.out_bf16_choff = outb.data_ptr(), 32, 0
    d.num_kblocks = 6
    for i in range(1, 6):
        d.kblocks[i] = d.kblocks[0]
    l.esr_debug_set_profile_buffer(C.c_void_p(prof.data_ptr()))
    for it in range(3):
        prof.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run_conv(d, 'tc'); e1.record(); torch.cuda.synchronize()
    q = prof.cpu().double(); lead = q[0::2].mean(0)
    print('buffer channels %3d: %.1f us; producer per kb total %.0f wait_empty %.0f; mma busy %.0f wait_full %.0f wait_acc %.0f' % (
        bc, e0.elapsed_time(e1) * 1e3, lead[0] / lead[2], lead[1] / lead[2], lead[3] - lead[4] - lead[5] - lead[9], lead[5], lead[4]))
