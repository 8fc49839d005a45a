"""Per-role cycle breakdown of the fused growth-conv launch (debug aid, GPU box; library built with -DESR_PROFILE_ROLES)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, '.')
os.environ["ESR_FUSE_RDB"] = "1"
from esr_b200 import _capi as capi, synth
from oracle.cem_ops import concat_latent
from tests.test_gpu_net import build_product_G
dev = torch.device('cuda', 0)
chunk = sys.argv[1] if len(sys.argv) > 1 else "8"
os.environ["ESR_RDB_CHUNK"] = chunk
l = capi.lib()
l.esr_debug_set_profile_buffer.argtypes = [C.c_void_p]
prof = torch.zeros(148 * 16 + 32 * 8, dtype=torch.int64, device=dev)
l.esr_debug_set_profile_buffer(C.c_void_p(prof.data_ptr()))
netG = build_product_G(dev, 1, 'all_layers_HR_downscaled', synth.make_weights('default', seed=0, nb=1))
G = netG.generated_image_model
lr, z = synth.make_inputs(16, 128, 128, seed=0)
x = concat_latent(lr, z).to(dev)
with torch.no_grad():
    for _ in range(3):
        netG(x)
    torch.cuda.synchronize()
    plan = list(G._plans.values())[-1]
    d = [o for o in plan.ops if isinstance(o, capi.RdbGrowthDesc)][0]
    for it in range(3):
        prof.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); capi.check(l.esr_rdb_growth_tc(C.byref(d), capi.stream_ptr())); e1.record(); torch.cuda.synchronize()
tr = prof[148 * 16:].cpu().view(32, 8).double()
q = prof[:148 * 16].view(148, 16).cpu().double()
lead = q[0::2].mean(0); allc = q.mean(0)
print('fused growth launch, chunk %s: %.1f us event-timed (NOTE: standalone relaunch re-uses dirty counters: dependencies trivially met)' % (chunk, e0.elapsed_time(e1) * 1e3))
print('  producer: total %.0f wait_empty %.0f poll %.0f respins %.0f items %.0f -> per item total %.0f wait_empty %.0f poll %.0f' % (
    allc[0], allc[1], allc[6], allc[10], allc[2], allc[0] / allc[2], allc[1] / allc[2], allc[6] / allc[2]))
print('  producer A detail per item: issue(expect_tx+tma) %.0f  syncwarp %.0f  poll-next %.0f' % (allc[12] / allc[2], allc[13] / allc[2], allc[14] / allc[2]))
print('  mma: total %.0f wait_acc_empty %.0f wait_full %.0f wait_w %.0f -> busy %.0f (per item %.0f)' % (
    lead[3], lead[4], lead[5], lead[9], lead[3] - lead[4] - lead[5] - lead[9], (lead[3] - lead[4] - lead[5] - lead[9]) / allc[2]))
print('  epi(warp2): total %.0f wait_acc_full %.0f items %.0f -> per item total %.0f wait %.0f work %.0f' % (
    allc[7], allc[8], allc[11], allc[7] / allc[11], allc[8] / allc[11], (allc[7] - allc[8]) / allc[11]))

t0 = tr[0, 0]
print('  trace of cluster 0 (us from its first item): item | producer: item start, loads issued | mma: acc free, first tile landed, last tile landed | epilogue: acc ready, stores issued')
for i in range(20):
    print('   %2d | %6.2f %6.2f | %6.2f %6.2f %6.2f | %6.2f %6.2f' % ((i,) + tuple(float((tr[i, k] - t0) / 1e3) for k in range(7))))
