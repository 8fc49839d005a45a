"""Shared helpers of the GPU parity tests: single-conv cases built through the same PackedConv /
ConvDesc path the engine uses, checked against the CPU oracle's conv (F.conv2d, fp32)."""
import ctypes as C

import numpy as np
import torch
import torch.nn.functional as F

from esr_b200 import _capi as capi
from esr_b200.engine import PackedConv, DY_ALL, DY_CENTRE


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def psnr(a, b):
    return float(10 * torch.log10(1.0 / ((a.double() - b.double()) ** 2).mean()))


def plain_conv_case(dev, B, H, W, cin, cout, seed=0, buf_channels=None, chan0=0, precise=False, pair=False, cout_tile=None):
    """Random conv with `cin` (multiple of 32) NHWC bf16 input channels living at [chan0, chan0+cin) of
    a buffer with `buf_channels` channels.  precise: input given as hi/lo pair (lo at +64, cin must be 64)."""
    g = torch.Generator().manual_seed(seed)
    w = (torch.rand(cout, cin, 3, 3, generator=g) - 0.5) * (2.0 / np.sqrt(cin * 9))
    b = torch.rand(cout, generator=g) - 0.5
    x = torch.rand(B, cin, H, W, generator=g) * 2 - 1
    bc = buf_channels or (128 if precise else cin)
    buf = torch.zeros(B, H, W, bc)
    hi = bf16_round(x)
    buf[..., chan0:chan0 + cin] = hi.permute(0, 2, 3, 1)
    if precise:
        assert cin == 64 and chan0 == 0
        buf[..., 64:128] = bf16_round(x - hi).permute(0, 2, 3, 1)
    kb, slots = [], []
    terms = ((0, 0), (64, 0), (0, 1)) if precise else ((chan0, 0),)
    for base, wterm in terms:
        for c0 in range(0, cin, 32):
            kb.append((0, base + c0, DY_ALL, 0b11))
            slots += [(c0 + k, -1, wterm) for k in range(32)]
    ct = cout_tile or (16 if cout <= 16 else 32)
    pc = PackedConv("test", cout, kb, slots, [(co, -1) for co in range(cout)], ct, pair=pair)
    wd, bd = w.to(dev).contiguous(), b.to(dev).contiguous()
    pc.pack(wd, bd, 0, cin * 9, 9, 3, 1)
    torch.cuda.synchronize()
    x_eff = x if precise else hi                      # precise mode reconstructs x to ~2^-16
    w_eff = w if precise else bf16_round(w)
    ref = F.conv2d(x_eff, w_eff, b, padding=1)
    return dict(pc=pc, buf=buf.to(dev).to(torch.bfloat16).contiguous(), w=w, b=b, x=x, ref=ref, keep=(wd, bd))


def conv_desc(pc, B, H, W, src0, src1=None):
    d = capi.ConvDesc()
    d.B, d.H, d.W = B, H, W
    d.src[0].ptr, d.src[0].channels = src0.data_ptr(), src0.shape[-1]
    if src1 is not None:
        d.src[1].ptr, d.src[1].channels = src1.data_ptr(), src1.shape[-1]
    d.cout_tile, d.cout_tiles, d.num_kblocks, d.pair = pc.cout_tile, pc.cout_tiles, pc.nkb, pc.pair
    for i in range(pc.nkb):
        d.kblocks[i] = pc.kblocks[i]
    d.wpack, d.w_tile_bytes, d.bias = pc.wpack.data_ptr(), pc.w_tile_bytes, pc.bias.data_ptr()
    d.slope, d.alpha, d.beta, d.up, d.out_bf16_scale, d.out_bf16_lo_choff = 0.2, 1.0, 1.0, 1, 1.0, -1
    return d


def run_conv(d, impl):
    fn = capi.lib().esr_conv3x3_tc if impl == "tc" else capi.lib().esr_conv3x3_simt
    capi.check(fn(C.byref(d), capi.stream_ptr()))
    torch.cuda.synchronize()
