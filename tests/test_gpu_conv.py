"""conv3x3 kernels (tcgen05 and the SIMT cross-check) against the CPU oracle conv, through the C ABI."""
import pytest
import torch
import torch.nn.functional as F

from esr_b200 import _capi as capi
from tests.helpers import plain_conv_case, conv_desc, run_conv, bf16_round

pytestmark = pytest.mark.gpu
TOL = 2e-3   # bf16 operands are exact in both paths; only fp32 accumulation order differs


def _nchw(t_nhwc):
    return t_nhwc.float().permute(0, 3, 1, 2).cpu()


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("B,H,W,cin,cout", [(1, 8, 30, 32, 32), (2, 20, 20, 64, 32), (1, 13, 70, 96, 32),
                                            (1, 9, 33, 192, 64), (1, 40, 61, 64, 3), (2, 17, 95, 160, 32)])
def test_plain_conv_f32_out(cuda_device, impl, B, H, W, cin, cout):
    c = plain_conv_case(cuda_device, B, H, W, cin, cout, seed=cin + cout + W)
    pc = c["pc"]
    out = torch.full((B, H, W, pc.cout_tiles * pc.cout_tile), 7.0, device=cuda_device)
    d = conv_desc(pc, B, H, W, c["buf"])
    d.out_f32, d.out_f32_stride = out.data_ptr(), out.shape[-1]
    run_conv(d, impl)
    got = _nchw(out)[:, :cout]
    err = (got - c["ref"]).abs().max().item()
    assert err < TOL, "max abs err %g" % err


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("B,H,W,cin,cout,ct", [(1, 8, 30, 32, 32, 32), (2, 20, 20, 64, 32, 32), (1, 13, 70, 96, 64, 32),
                                               (1, 9, 33, 192, 64, 64), (3, 21, 95, 160, 32, 32), (2, 23, 61, 64, 64, 64),
                                               (1, 4, 30, 32, 64, 64), (1, 37, 151, 64, 128, 64),
                                               # several tiles per cluster (ring / accumulator-stage reuse) and
                                               # more K blocks per tile than ring stages
                                               (8, 70, 95, 224, 64, 64), (8, 70, 95, 256, 32, 32), (9, 41, 61, 64, 64, 32)])
def test_pair_conv_f32_out(cuda_device, impl, B, H, W, cin, cout, ct):
    """cta_group::2 kernel (CTA pairs, M = 256) and the pair weight layout: odd tile counts, one or two cout
    tiles, N = 96 (two bands per CTA) and N = 192 (one band per CTA)."""
    c = plain_conv_case(cuda_device, B, H, W, cin, cout, seed=cin + cout + W, pair=True, cout_tile=ct)
    pc = c["pc"]
    out = torch.full((B, H, W, pc.cout_tiles * pc.cout_tile), 7.0, device=cuda_device)
    d = conv_desc(pc, B, H, W, c["buf"])
    d.out_f32, d.out_f32_stride = out.data_ptr(), out.shape[-1]
    run_conv(d, impl)
    got = _nchw(out)[:, :cout]
    err = (got - c["ref"]).abs().max().item()
    assert err < TOL, "max abs err %g" % err


@pytest.mark.parametrize("ct", [32, 64])
def test_pair_residual_epilogue(cuda_device, ct):
    """conv5-like launch in pair mode: v = 0.2*(0.2*conv + res1) + res2 -> blocked fp32 trunk + bf16 slice."""
    B, H, W = 2, 12, 37
    c = plain_conv_case(cuda_device, B, H, W, 192, 64, seed=9, pair=True, cout_tile=ct)
    g = torch.Generator().manual_seed(1)
    r1, r2 = torch.rand(B, H, W, 64, generator=g), torch.rand(B, H, W, 64, generator=g)
    blk = lambda t: t.view(B, H, W, 8, 8).permute(0, 3, 1, 2, 4).contiguous()       # [B, C/8, H, W, 8]
    r1d, r2d = blk(r1).to(cuda_device), blk(r2).to(cuda_device)
    out32 = torch.zeros(B, 8, H, W, 8, device=cuda_device)
    outb = torch.zeros(B, H, W, 192, device=cuda_device, dtype=torch.bfloat16)
    d = conv_desc(c["pc"], B, H, W, c["buf"])
    d.flags = capi.EPI_RES1 | capi.EPI_RES2 | capi.EPI_F32_BLOCKED
    d.alpha, d.beta = 0.2, 0.2
    d.res1, d.res1_stride, d.res2, d.res2_stride = r1d.data_ptr(), 64, r2d.data_ptr(), 64
    d.out_f32, d.out_f32_stride = out32.data_ptr(), 64
    d.out_bf16, d.out_bf16_stride, d.out_bf16_choff = outb.data_ptr(), 192, 0
    run_conv(d, "tc")
    ref = 0.2 * (0.2 * c["ref"] + r1.permute(0, 3, 1, 2)) + r2.permute(0, 3, 1, 2)
    got = out32.cpu().permute(0, 1, 4, 2, 3).reshape(B, 64, H, W)
    assert (got - ref).abs().max().item() < TOL
    assert (_nchw(outb)[:, :64] - bf16_round(ref)).abs().max().item() < 1e-2


@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_dense_block_slice_write_and_lrelu(cuda_device, impl):
    """conv reads channels [0,96) of a 192-channel buffer and writes LeakyReLU(out) as bf16 into [96,128)."""
    B, H, W = 1, 21, 45
    c = plain_conv_case(cuda_device, B, H, W, 96, 32, seed=5, buf_channels=192)
    buf = c["buf"]
    d = conv_desc(c["pc"], B, H, W, buf)
    d.flags = capi.EPI_LRELU
    d.out_bf16, d.out_bf16_stride, d.out_bf16_choff = buf.data_ptr(), 192, 96
    run_conv(d, impl)
    got = _nchw(buf)[:, 96:128]
    ref = bf16_round(F.leaky_relu(c["ref"], 0.2))
    assert (got - ref).abs().max().item() < 2e-2          # one bf16 ulp at |v|<=2
    assert ((got - ref).abs() > 1e-6).float().mean().item() < 0.02
    assert torch.equal(_nchw(buf)[:, :96], _nchw(c["buf"])[:, :96])


@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_residual_epilogue_and_split_output(cuda_device, impl):
    """v = 0.2*(0.2*conv + res1) + res2 -> fp32 trunk, bf16 hi/lo pair (RDB3 / last-RDB epilogue)."""
    B, H, W = 2, 12, 37
    c = plain_conv_case(cuda_device, B, H, W, 192, 64, seed=9)
    g = torch.Generator().manual_seed(1)
    r1, r2 = torch.rand(B, H, W, 64, generator=g), torch.rand(B, H, W, 64, generator=g)
    r1d, r2d = r1.to(cuda_device), r2.to(cuda_device)
    out32 = torch.zeros(B, H, W, 64, device=cuda_device)
    outb = torch.zeros(B, H, W, 192, device=cuda_device, dtype=torch.bfloat16)
    d = conv_desc(c["pc"], B, H, W, c["buf"])
    d.flags = capi.EPI_RES1 | capi.EPI_RES2
    d.alpha, d.beta = 0.2, 0.2
    d.res1, d.res1_stride, d.res2, d.res2_stride = r1d.data_ptr(), 64, r2d.data_ptr(), 64
    d.out_f32, d.out_f32_stride = out32.data_ptr(), 64
    d.out_bf16, d.out_bf16_stride, d.out_bf16_choff, d.out_bf16_lo_choff = outb.data_ptr(), 192, 0, 64
    run_conv(d, impl)
    ref = 0.2 * (0.2 * c["ref"] + r1.permute(0, 3, 1, 2)) + r2.permute(0, 3, 1, 2)
    assert (_nchw(out32) - ref).abs().max().item() < TOL
    recon = _nchw(outb)[:, :64] + _nchw(outb)[:, 64:128]
    assert (recon - ref).abs().max().item() < TOL


@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_split_bf16_input_and_up2_output(cuda_device, impl):
    """precise mode: hi/lo input pair, three MMA terms ~ fp32 conv; output replicated 2x2 (nearest)."""
    B, H, W = 1, 11, 40
    c = plain_conv_case(cuda_device, B, H, W, 64, 64, seed=3, precise=True)
    outb = torch.zeros(B, 2 * H, 2 * W, 128, device=cuda_device, dtype=torch.bfloat16)
    d = conv_desc(c["pc"], B, H, W, c["buf"])
    d.up = 2
    d.out_bf16, d.out_bf16_stride, d.out_bf16_choff, d.out_bf16_lo_choff = outb.data_ptr(), 128, 0, 64
    run_conv(d, impl)
    recon = _nchw(outb)[:, :64] + _nchw(outb)[:, 64:128]
    ref = F.interpolate(c["ref"], scale_factor=2, mode="nearest")
    assert (recon - ref).abs().max().item() < 2e-4


def test_nchw_output_cout3(cuda_device):
    B, H, W = 2, 19, 50
    c = plain_conv_case(cuda_device, B, H, W, 64, 3, seed=4)
    out = torch.zeros(B, 3, H, W, device=cuda_device)
    d = conv_desc(c["pc"], B, H, W, c["buf"])
    d.out_nchw, d.cout_real = out.data_ptr(), 3
    run_conv(d, "tc")
    assert (out.cpu() - c["ref"]).abs().max().item() < TOL


@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_dgrad_packing_matches_conv_transpose(cuda_device, impl):
    """dgrad = the same kernel on the transposed + flipped weight view: d(input) of a 67->32 conv, with the
    first 3 input channels (latent) reported as 9 row-expanded rows (centre tap, fixed filter row)."""
    from esr_b200.engine import PackedConv, DY_ALL
    from tests.helpers import bf16_round
    B, H, W, cin, cout, nz = 1, 14, 37, 67, 32, 3
    g = torch.Generator().manual_seed(11)
    w = (torch.rand(cout, cin, 3, 3, generator=g) - 0.5) * 0.2
    gy = bf16_round(torch.randn(B, cout, H, W, generator=g))
    rows = [(nz + c, -1) for c in range(64)] + [(c, 2 - dy) for dy in range(3) for c in range(nz)] + [(-1, -1)] * 23
    pc = PackedConv("dgrad", 96, [(0, 0, DY_ALL, 0b11)], [(k, -1, 0) for k in range(32)], rows, 32)
    wd = w.to(cuda_device).contiguous()
    pc.pack(wd, None, 8, 9, cin * 9, -3, -1)
    src = gy.permute(0, 2, 3, 1).contiguous().to(cuda_device).to(torch.bfloat16)
    out = torch.zeros(B, H, W, 96, device=cuda_device)
    d = conv_desc(pc, B, H, W, src)
    d.out_f32, d.out_f32_stride = out.data_ptr(), 96
    run_conv(d, impl)
    got = _nchw(out)
    wq = bf16_round(w)
    ref = F.conv_transpose2d(gy, wq, padding=1)                       # [B, 67, H, W]
    assert (got[:, :64] - ref[:, nz:]).abs().max().item() < TOL
    # latent rows: dE[(dy,c)][y, x] = sum_{co,dx} W[co,c,dy,dx] g[co, y, x-(dx-1)]  (no shift along y)
    for dy in range(3):
        for c in range(nz):
            k = torch.zeros(cout, 1, 1, 3)
            k[:, 0, 0, :] = wq[:, c, dy, :]
            e = F.conv_transpose2d(gy, k, padding=(0, 1))[:, 0]
            assert (got[:, 64 + dy * nz + c] - e).abs().max().item() < TOL
