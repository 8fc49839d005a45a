"""Host-side engine: packs RRDBNet weights into the conv kernels' shared-memory image and lays
out, per input geometry, the buffers and the recorded launch sequence of the G(+CEM) forward
(and data-gradient backward).  PyTorch is used for device memory and streams only; every
arithmetic step is a kernel of libesr_b200.so called through the C ABI.

Layer graph = codes/models/modules/architecture.py:102-175 with block.py:196-270 unrolled.
Data layout in HBM:
  * one NHWC bf16 "dense block" buffer [B,H,W,192] per RDB: channels 0..63 = block input x0,
    64+32i.. = growth x_{i+1}; conv i reads channels [0,64+32i) and writes its 32 outputs into
    its own slice, so torch.cat never materialises;
  * the residual trunk (RDB / RRDB / shortcut sums) stays fp32 NHWC [B,H,W,64];
  * the 3-channel latent is expanded once per resolution over the filter rows into a 32-channel
    bf16 tensor shared by all 346+2 convs that take it;
  * the six convs outside the residual-scaled trunk (first, LR_conv, 2 upconvs, 2 HR convs) carry
    most of the operand-rounding error (oracle experiment in DESIGN.md: all-bf16 55 dB / 8e-3 max,
    bf16 trunk + fp16 outer 72 dB / 1e-3) and 7.8 % of the FLOPs.  outer_mode "f16" (default) runs them
    with fp16 operands (one MMA term, fp32 accumulate); "split" keeps bf16 hi+lo pairs (three terms,
    ~2^-16 relative); "bf16" is the plain single-term bf16 form.
"""
import ctypes as C
import math
import os

import numpy as np
import torch

from . import _capi as capi
from ._capi import KBlock, ConvDesc, WRow, WSlot, XSlot, RdbGrowthDesc

DY_ALL, DY_CENTRE = 0b111, 0b010
LEAN_LATENT = os.environ.get("ESR_LEAN_LATENT", "1") != "0"   # 0: load the latent rows as full 32-channel halo tiles (A/B timing)
NF, GC = 64, 32


def _struct_array_to_device(arr, ctype, device):
    """ctypes struct array -> uint8 CUDA tensor with the same bytes."""
    raw = bytes(memoryview(arr))
    t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
    return t


class PackedConv:
    """One logical 3x3 conv packed for the kernels: K-block list + swizzled bf16 weight image."""

    def __init__(self, name, cout, kblocks, slots, rows, cout_tile, pair=False):
        self.name, self.cout, self.cout_tile, self.pair = name, cout, cout_tile, int(bool(pair))
        self.cout_tiles = (cout + cout_tile - 1) // cout_tile
        self.nkb = len(kblocks)
        assert self.nkb <= capi.MAX_KBLOCKS, name
        self.kblocks = (KBlock * capi.MAX_KBLOCKS)()
        # lean latent blocks (esr_kblock.half): every block of source 1 is centre tap + one 16-channel slice; pair only
        src1 = [kbt for kbt in kblocks if kbt[0] == 1]
        lean = bool(self.pair and LEAN_LATENT and src1 and all(k[2] == DY_CENTRE and k[3] in (1, 2) for k in src1))
        for i, kbt in enumerate(kblocks):
            src, chan, dy_mask, slice_mask = kbt[:4]
            self.kblocks[i].src, self.kblocks[i].chan = src, chan
            self.kblocks[i].dy_mask, self.kblocks[i].slice_mask = dy_mask, slice_mask
            self.kblocks[i].half = 1 if (lean and src == 1) else 0
        wtb = C.c_uint32(0)
        total = capi.lib().esr_pack_layout(cout_tile, self.cout_tiles, self.pair, self.nkb, self.kblocks, C.byref(wtb))
        if total < 0:
            capi.check(int(total))
        self.w_tile_bytes, self.total_bytes = wtb.value, int(total)
        self.slots = (WSlot * (self.nkb * 32))()
        for i, (idx, ky, term) in enumerate(slots):
            self.slots[i].idx, self.slots[i].ky, self.slots[i].term = idx, ky, term
        nrows = self.cout_tiles * cout_tile
        self.rows = (WRow * nrows)()
        for i in range(nrows):
            idx, ky = rows[i] if i < len(rows) else (-1, -1)
            self.rows[i].idx, self.rows[i].ky = idx, ky
        self.wpack = self.bias = None

    def pack(self, weight, bias, off, s_row, s_slot, s_ky, s_kx, table=None):
        """Packs this conv's weight image.  With `table` (a list) nothing is launched: the entry of the batched pack
        table (esr_pack_entry_fill) is appended instead and PackTable.run() packs every conv of the table in one launch."""
        dev = weight.device
        if self.wpack is None or self.wpack.device != dev:     # re-packing after a weight update reuses the buffers: the recorded
            self.wpack = torch.empty(self.total_bytes, dtype=torch.uint8, device=dev)      # launch sequences point at them
            self.bias = torch.empty(self.cout_tiles * self.cout_tile, dtype=torch.float32, device=dev)
        keep = getattr(self, "_keep", None)
        if keep is None or keep[0].device != dev:              # the row / slot tables never change: uploaded once (a training
            keep = (_struct_array_to_device(self.rows, WRow, dev),           # step re-packs all 351 convs after every update)
                    _struct_array_to_device(self.slots, WSlot, dev))
        rows_d, slots_d = keep
        self._keep = keep
        args = (capi.ptr(weight), off, s_row, s_slot, s_ky, s_kx, capi.ptr(bias), self.cout_tile, self.cout_tiles,
                self.pair, self.nkb, self.kblocks, self.w_tile_bytes, capi.ptr(rows_d), capi.ptr(slots_d), capi.ptr(self.wpack),
                capi.ptr(self.bias))
        if table is not None:
            entry = (C.c_uint8 * capi.lib().esr_pack_entry_bytes())()
            capi.check(capi.lib().esr_pack_entry_fill(entry, *args))
            table.append(bytes(entry))
            return
        capi.check(capi.lib().esr_pack_conv_weights(*args, capi.stream_ptr()))


class PackTable:
    """Every conv of a network packed by ONE launch (esr_pack_table_run): entries are built once per set of source
    pointers, re-run after every weight update (training: 351 + ~420 images per step)."""

    def __init__(self):
        self.key, self.dev, self.n = None, None, 0

    def build(self, key, entries, device):
        raw = b"".join(entries)
        self.dev = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
        self.key, self.n = key, len(entries)

    def run(self):
        capi.check(capi.lib().esr_pack_table_run(capi.ptr(self.dev), self.n, capi.stream_ptr()))


def expand_slots(nvals_channels, precise, second="lo"):
    """Slot table of a row-expanded small-channel tensor.

    values v = (dy, c) for dy in 0..2, c in range(nvals_channels).
    precise == "f16": [fp16(v)...] padded to a multiple of 32 (one term)
    precise is True : [hi(v)... | lo(v)... | hi(v)...] padded to a multiple of 32 (three split-bf16 terms)
    else            : [hi(v)..., pad to 16 | second(v)..., pad to 16] where second is the bf16 residue ("lo") or
                      fp16(v) ("f16"); the trunk convs read slice 0 only, the outer convs slice 1 (f16) or both
    Returns (xslots for esr_expand_rows, wslots template [(value index, dy, term)] per slot).
    """
    vals = [(dy, c) for dy in range(3) for c in range(nvals_channels)]
    V = len(vals)
    xs, ws = [], []
    if precise == "f16":
        for (dy, c) in vals:
            xs.append((c, dy - 1, 2))
            ws.append((c, dy, 2))
        while len(xs) % 32:
            xs.append((-1, 0, 0))
            ws.append((-1, -1, 0))
    elif precise:
        n = ((3 * V + 31) // 32) * 32
        for part, (xterm, wterm) in enumerate(((0, 0), (1, 0), (0, 1))):
            for (dy, c) in vals:
                xs.append((c, dy - 1, xterm))
                ws.append((c, dy, wterm))
        while len(xs) < n:
            xs.append((-1, 0, 0))
            ws.append((-1, -1, 0))
    else:
        assert V <= 16
        for xterm in (0, 2 if second == "f16" else 1):
            for (dy, c) in vals:
                xs.append((c, dy - 1, xterm))
                ws.append((c, dy, 2 if xterm == 2 else 0))
            while len(xs) % 16:
                xs.append((-1, 0, 0))
                ws.append((-1, -1, 0))
    return xs, ws


def _xslot_array(xs):
    arr = (XSlot * len(xs))()
    for i, (c, dy, term) in enumerate(xs):
        arr[i].c, arr[i].dy, arr[i].term = c, dy, term
    return arr


class GEngine:
    """Packed weights of one RRDBNet; geometry-independent."""

    def __init__(self, nb, nz_in, all_layers, out_nc=3, in_nc=3, upscale=4, precise_outer=True, pair=True,
                 outer_mode=None, z_rearranged=0):
        if upscale not in (2, 4):
            # x3: the reference's own RRDBNet cannot be constructed for upscale=3 (architecture.py:144 concatenates a
            # list with the nn.Sequential its x3 upsampler is: TypeError), so there is nothing to be a drop-in for
            raise NotImplementedError("upscale %d: only x2 / x4 (nearest x2 upconv stages) are built" % upscale)
        if in_nc != 3 or out_nc > 16:
            raise NotImplementedError("in_nc must be 3 and out_nc <= 16")
        self.nb, self.nz_in, self.all_layers = nb, nz_in, all_layers
        # first_layer_HR_rearranged (architecture.py:109-110,159): `z_rearranged` = Cz * sf^2 extra input channels of the
        # first conv, given at LR resolution; that conv then has the shape of the outer convs (a 64-channel 16-bit tensor
        # holding Z, zero padded, + the 3 image channels as the row-expanded small block)
        self.z_rearranged = int(z_rearranged)
        if self.z_rearranged and (nz_in or self.z_rearranged > NF):
            raise NotImplementedError("HR_rearranged latent: first_layer only, at most %d rearranged channels" % NF)
        self.nz = nz_in if all_layers else 0          # latent channels concatenated to every later conv
        self.out_nc, self.upscale = out_nc, upscale
        self.n_up = int(math.log2(upscale))
        self.precise = precise_outer
        # backward: the dgrads of the outer convs take plain bf16 gradient operands like the 345 trunk dgrads (measured
        # against the oracle's autograd at nb=23: relative error 0.0285 vs 0.0280 with three-term split-bf16 operands,
        # 4 % faster per Z iteration); ESR_BWD_PRECISE=1 restores the split operands
        self.precise_bwd = precise_outer and os.environ.get("ESR_BWD_PRECISE", "0") == "1"
        self.outer_mode = outer_mode or os.environ.get("ESR_OUTER_MODE") or ("f16" if precise_outer else "bf16")
        assert self.outer_mode in ("f16", "split", "bf16")
        # CTA-pair (cta_group::2) kernels wherever cout >= 32; ESR_PAIR=0 keeps the single-CTA kernel (A/B timing)
        self.pair = pair and os.environ.get("ESR_PAIR", "1") != "0"
        self.convs = {}
        self.version = None
        self._build_specs()

    # ------------------------------------------------------------------ specs
    def _main_blocks(self, nch, mode, src=0):
        """K blocks + weight slots for `nch` feature channels at buffer channels [0,nch) (bf16 hi, or fp16 in mode
        "f16") and, in mode "split", their bf16 residues at [64,64+nch)."""
        kb, sl = [], []
        terms = {"split": ((0, 0), (64, 0), (0, 1)), "f16": ((0, 2),)}.get(mode, ((0, 0),))
        for base, wterm in terms:
            for c0 in range(0, nch, 32):
                kb.append((src, base + c0, DY_ALL, 0b11))
                sl += [(self.nz + c0 + k, -1, wterm) for k in range(32)]
        return kb, sl

    def _latent_blocks(self, mode, src=1):
        if self.nz == 0:
            return [], []
        _, ws = expand_slots(self.nz, precise=False, second=self.lat_second)   # E_lat layout: [hi | lo or f16], 16 + 16
        kb, sl = [], []
        if mode == "f16":
            kb.append((src, 0, DY_CENTRE, 0b10))       # the fp16 copy of the latent rows lives in slice 1
            sl += list(ws)
        elif mode == "split":
            kb.append((src, 0, DY_CENTRE, 0b11))
            sl += list(ws)                             # A_hi*W_hi and A_lo*W_hi
            kb.append((src, 0, DY_CENTRE, 0b01))
            sl += [(i, ky, 1) for (i, ky, _) in ws]    # A_hi*W_lo
        else:
            kb.append((src, 0, DY_CENTRE, 0b01))
            sl += list(ws)
        return kb, sl

    def _build_specs(self):
        p = self.outer_mode
        self.lat_second = "f16" if p == "f16" else "lo"
        # first conv: every input is row-expanded (E_fea), centre tap only
        self.fea_xslots, fea_ws = expand_slots(self.nz_in + 3, precise="f16" if p == "f16" else True)
        if self.z_rearranged:
            if p == "split":
                raise NotImplementedError("HR_rearranged latent with outer_mode 'split'")
            zr = self.z_rearranged
            kb, sl = self._main_blocks(NF, p)                  # weight input channel c < zr: Z channel c of the 16-bit tensor
            sl = [(i if 0 <= i < zr else -1, ky, t) for (i, ky, t) in sl]
            _, ws = expand_slots(3, precise=False, second=self.lat_second)     # the image: weight input channels zr .. zr + 2
            ws = [(zr + i if i >= 0 else i, ky, t) for (i, ky, t) in ws]
            kb.append((1, 0, DY_CENTRE, 0b10 if p == "f16" else 0b01))
            self._add("model.0", NF, kb, sl + ws, 32)
            self.lr_xslots, _ = expand_slots(3, precise=False, second=self.lat_second)
        else:
            kb = [(0, c0, DY_CENTRE, 0b11) for c0 in range(0, len(fea_ws), 32)]
            self._add("model.0", NF, kb, fea_ws, 32)
        self.lat_xslots, _ = expand_slots(max(self.nz, 1), precise=False, second=self.lat_second)
        for r in range(self.nb):
            for d in (1, 2, 3):
                for i in range(5):
                    kb, sl = self._main_blocks(NF + GC * i, "bf16")
                    kb2, sl2 = self._latent_blocks("bf16")
                    self._add("model.1.sub.%d.RDB%d.convs.%d.0" % (r, d, i), GC if i < 4 else NF, kb + kb2, sl + sl2, 32)
        outer = ["model.1.sub.%d" % self.nb] + ["model.%d.1" % (2 + u) for u in range(self.n_up)] + \
                ["model.%d" % (2 + self.n_up), "model.%d" % (4 + self.n_up)]
        self.outer_names = outer
        self.upconv_names = set(outer[1:1 + self.n_up])
        for name in outer:
            has_lat = name not in self.upconv_names    # upconvs take no latent (architecture.py:164-171)
            kb, sl = self._main_blocks(NF, p)
            if not has_lat:
                sl = [(i - self.nz if i >= 0 else i, ky, t) for (i, ky, t) in sl]
                kb2, sl2 = [], []
            else:
                kb2, sl2 = self._latent_blocks(p)
            last = name == outer[-1]
            self._add(name, self.out_nc if last else NF, kb + kb2, sl + sl2, 16 if last else 32)
        self.f16_convs = set(["model.0"] + outer) if p == "f16" else set()

    def _add(self, name, cout, kblocks, slots, cout_tile):
        rows = [(co, -1) for co in range(cout)]
        pair = self.pair and cout_tile == 32
        if pair and cout % 64 == 0:
            cout_tile = 64                             # N = 192: the activation tile is read once for all 64 channels
        self.convs[name] = PackedConv(name, cout, kblocks, slots, rows, cout_tile, pair=pair)

    # ---------------------------------------------------------------- packing
    def pack(self, params):
        """params: dict name -> (weight OIHW f32 CUDA contiguous, bias f32 CUDA)."""
        key = tuple((params[name][0].data_ptr(), params[name][1].data_ptr()) for name in self.convs)
        tab = getattr(self, "_pack_table", None)
        if tab is None or tab.key != key:
            entries = []
            for name, pc in self.convs.items():
                w, b = params[name]
                cin = w.shape[1]
                pc.pack(w, b, 0, cin * 9, 9, 3, 1, table=entries)
                pc.cin = cin
            tab = self._pack_table = PackTable()
            tab.build(key, entries, next(iter(params.values()))[0].device)
        tab.run()

    def flops_per_lr_pixel(self):
        """Algorithmic MACs*2 of the reference network per (padded) LR pixel."""
        total = 0
        for name, pc in self.convs.items():
            res = 1
            if name in self.outer_names[1:]:
                u = self.outer_names.index(name)
                res = 4 ** min(u, self.n_up)
            total += 2 * 9 * pc.cin * pc.cout * res
        return total


class GPlan:
    """Buffers + recorded conv sequence for one geometry (B, h, w, margin)."""

    def __init__(self, eng, B, h, w, m, device, with_cem_input=True, keep_activations=False, use_simt=False):
        self.eng, self.B, self.h, self.w, self.m = eng, B, h, w, m
        self.device = device
        self.graphed = {}                      # (cem filters, margin) -> rrdbnet._GraphedStep (dies with the plan)
        sf = eng.upscale
        hp, wp = h + 2 * m, w + 2 * m
        self.hp, self.wp, self.sf = hp, wp, sf
        f32 = dict(dtype=torch.float32, device=device)
        bf = dict(dtype=torch.bfloat16, device=device)
        nzi, nz = eng.nz_in, eng.nz
        self.lr_pad = torch.empty(B, 3, hp, wp, **f32)
        self.fea_in = torch.empty(B, nzi + 3, hp, wp, **f32)
        self.z_hr = torch.empty(B, nz, sf * hp, sf * wp, **f32) if nz else None
        self.z_lr = torch.empty(B, nz, hp, wp, **f32) if nz else None
        self.E_fea = torch.empty(B, hp, wp, len(eng.fea_xslots), **bf)
        if eng.z_rearranged:                                   # Z as a 64-channel 16-bit tensor (zero padded) + the image rows
            self.Zbuf = torch.zeros(B, hp, wp, NF, **bf)
            self.E_lr = torch.empty(B, hp, wp, 32, **bf)
        self.E_lat = torch.empty(B, hp, wp, 32, **bf) if nz else None
        self.E_lath = torch.empty(B, sf * hp, sf * wp, 32, **bf) if nz else None
        n_rdb = 3 * eng.nb
        nbuf = n_rdb + 1 if keep_activations else 2
        self.bufs = [torch.empty(B, hp, wp, 192, **bf) for _ in range(nbuf)]
        self.T_fea = torch.empty(B, hp, wp, NF, **f32)
        self.R = [torch.empty(B, hp, wp, NF, **f32) for _ in range(2)]
        # Opt-in (ESR_TRUNK_HILO=1): carry the RDB -> RDB residual inside an RRDB as bf16 hi (channels 0..63 of the next
        # dense-block buffer, written anyway) + bf16 lo instead of fp32.  Saves 18 % of conv 4's HBM bytes but its
        # pixel-strided 32-byte reads are slower than the blocked fp32 layout's 1 KiB runs: measured 16.8 vs 16.0 ms/step.
        self.trunk_hilo = os.environ.get("ESR_TRUNK_HILO", "0") == "1"
        self.T = [torch.empty(B, hp, wp, NF, **(bf if self.trunk_hilo else f32)) for _ in range(2)]
        res = [(2 ** (u + 1)) for u in range(eng.n_up)]       # 2, 4
        oc = 128 if eng.outer_mode == "split" else 64        # split mode stores bf16 hi | lo pairs
        self.U = [torch.empty(B, r * hp, r * wp, oc, **bf) for r in res]   # nearest-upsampled inputs of the upconvs
        H4, W4 = sf * hp, sf * wp
        self.V1 = torch.empty(B, H4, W4, oc, **bf)
        self.V2 = torch.empty(B, H4, W4, oc, **bf) if (keep_activations or eng.n_up < 2) else self.U[-1]
        self.y = torch.empty(B, eng.out_nc, H4, W4, **f32)
        self.seq = capi.lib().esr_seq_create()
        self.descs = []
        # Fused growth convs (conv 0..3 of every RDB in one persistent launch with tile-level dependencies, csrc/
        # conv3x3_tc2.cu): bit-identical to the separate launches, 78 % less DRAM traffic for those convs at 4-8 images per
        # chunk, 207 launches fewer.  With all four weight images resident only 4 A-ring stages fit, which costs 3 % at
        # config 2 - but small plans are bound by the per-launch fixed cost and tile-round quantisation, not
        # by the ring: there the fused launch wins (config 3: forward + backward 10.6 vs 10.9 ms in round 1).  Default:
        # fused when the plan has at most ESR_FUSE_RDB_MAX_PIXELS padded LR pixels (120 k: config 3 = 76 k, two images of
        # config 2 = 44 k, config 2 itself = 350 k); ESR_FUSE_RDB=1 / 0 forces it on / off.
        # (Round 1 kept plans below 32 x 32 padded pixels on separate launches because a 12 x 14 plan faulted: the item
        # decode divided by 1 through a magic number that had wrapped to 0 - fixed in csrc/conv3x3_tc2.cu: fast_div -
        # so every plan size takes the fused launch now; tests/test_gpu_net.py covers 12x14, 22x24, 33x35.)
        # ... and only when a layer has at least two rounds of tile pairs (2 x 74 clusters): below that the clusters run
        # into the next layer's items and stall on their neighbours (tools/fuse_shapes.py on B200, forward graph, separate vs
        # fused launches: 1x64x64 = 33 pairs 1.78 / 1.89 ms).  Many small images need all of them in one chunk for the same
        # reason (csrc/conv3x3_tc2.cu: images per chunk; with chunks of 4 the fused launches were twice as slow on 16 x 32x32).
        force = os.environ.get("ESR_FUSE_RDB")
        pairs = B * ((((hp + 7) // 8) * ((wp + 29) // 30) + 1) // 2)
        small = B * hp * wp <= int(os.environ.get("ESR_FUSE_RDB_MAX_PIXELS", 120000)) and \
            pairs >= int(os.environ.get("ESR_FUSE_RDB_MIN_PAIRS", 90))
        self.fuse_rdb = eng.pair and not use_simt and (force == "1" or (force is None and small))
        self.rdb_flags = torch.zeros(int(capi.lib().esr_rdb_growth_flag_words(B, hp, wp)), dtype=torch.int32, device=device) \
            if self.fuse_rdb else None
        self._record_forward(use_simt)

    def __del__(self):
        try:
            if getattr(self, "seq", None):
                capi.lib().esr_seq_destroy(self.seq)
        except Exception:
            pass

    def buf(self, g):
        return self.bufs[g] if len(self.bufs) > 2 else self.bufs[g % 2]

    def _desc(self, name, H, W, src0, src1=None, flags=0, alpha=1.0, beta=1.0, res1=None, res2=None,
              out_bf16=None, out_choff=0, lo_choff=-1, up=1, out_f32=None, out_nchw=None, res1_hilo=None, out_lo=None):
        pc = self.eng.convs[name]
        d = ConvDesc()
        d.B, d.H, d.W = self.B, H, W
        d.src[0].ptr, d.src[0].channels = src0.data_ptr(), src0.shape[-1]
        if src1 is not None:
            d.src[1].ptr, d.src[1].channels = src1.data_ptr(), src1.shape[-1]
        d.cout_tile, d.cout_tiles, d.num_kblocks, d.pair = pc.cout_tile, pc.cout_tiles, pc.nkb, pc.pair
        for i in range(pc.nkb):
            d.kblocks[i] = pc.kblocks[i]
        d.wpack, d.w_tile_bytes, d.bias = pc.wpack.data_ptr(), pc.w_tile_bytes, pc.bias.data_ptr()
        d.flags, d.slope, d.alpha, d.beta = flags | capi.EPI_F32_BLOCKED, 0.2, alpha, beta   # trunk f32 = [B,8,H,W,8]
        if name in self.eng.f16_convs:
            d.flags |= capi.CONV_F16
        if res1 is not None:
            d.res1, d.res1_stride, d.res1_choff = res1.data_ptr(), res1.shape[-1], 0
            d.flags |= capi.EPI_RES1
        if res1_hilo is not None:                      # (hi tensor, lo tensor): bf16 NHWC, channels 0..63 of each
            hi_t, lo_t = res1_hilo
            d.res1_hi, d.res1_hi_stride, d.res1_hi_choff = hi_t.data_ptr(), hi_t.shape[-1], 0
            d.res1_lo, d.res1_lo_stride, d.res1_lo_choff = lo_t.data_ptr(), lo_t.shape[-1], 0
            d.flags |= capi.EPI_RES1 | capi.EPI_RES1_HILO
        if out_lo is not None:
            d.out_lo, d.out_lo_stride, d.out_lo_choff = out_lo.data_ptr(), out_lo.shape[-1], 0
        if res2 is not None:
            d.res2, d.res2_stride, d.res2_choff = res2.data_ptr(), res2.shape[-1], 0
            d.flags |= capi.EPI_RES2
        d.up, d.out_bf16_scale, d.out_bf16_lo_choff = up, 1.0, lo_choff
        if out_bf16 is not None:
            d.out_bf16, d.out_bf16_stride, d.out_bf16_choff = out_bf16.data_ptr(), out_bf16.shape[-1], out_choff
        if out_f32 is not None:
            d.out_f32, d.out_f32_stride, d.out_f32_choff = out_f32.data_ptr(), out_f32.shape[-1], 0
        if out_nchw is not None:
            d.out_nchw, d.cout_real = out_nchw.data_ptr(), pc.cout
        return d

    def _growth_desc(self, names, buf, lat, k, n, mode=0, mask=None, flags=None):
        """esr_rdb_growth_desc of launch k of n: the convs `names` read / write the dense-block buffer `buf`."""
        d = RdbGrowthDesc()
        d.B, d.H, d.W = self.B, self.hp, self.wp
        d.src[0].ptr, d.src[0].channels = buf.data_ptr(), buf.shape[-1]
        if lat is not None:
            d.src[1].ptr, d.src[1].channels = lat.data_ptr(), lat.shape[-1]
        d.num_layers = len(names)
        for l, (pc, choff) in enumerate(names):
            assert pc.pair and pc.cout_tile == 32 and pc.cout_tiles == 1 and pc.nkb <= capi.RDB_MAX_KBLOCKS
            ly = d.layers[l]
            ly.num_kblocks = pc.nkb
            for i in range(pc.nkb):
                ly.kblocks[i] = pc.kblocks[i]
            ly.wpack, ly.w_tile_bytes, ly.bias, ly.out_choff = pc.wpack.data_ptr(), pc.w_tile_bytes, pc.bias.data_ptr(), choff
        d.out, d.out_stride, d.mode, d.slope = buf.data_ptr(), buf.shape[-1], mode, 0.2
        if mask is not None:
            d.mask, d.mask_stride = mask.data_ptr(), mask.shape[-1]
        d.imgs_per_chunk = int(os.environ.get("ESR_RDB_CHUNK", 0))
        d.flags = (flags if flags is not None else self.rdb_flags).data_ptr()
        assert n >= 2
        use = lambda i: 2 if (i == n - 1 and n % 2 == 1) else i % 2       # consecutive launches never share a third,
        d.flags_use, d.flags_zero = use(k), use((k + 1) % n)               # including last -> first of the next replay
        return d

    def _record_forward(self, use_simt):
        eng, hp, wp = self.eng, self.hp, self.wp
        L = capi.EPI_LRELU
        self.ops = []                                      # ConvDesc | RdbGrowthDesc, in launch order
        add = lambda d: (self.descs.append(d), self.ops.append(d))
        n_rdb = 3 * eng.nb
        lo = 64 if eng.outer_mode == "split" else -1
        F16 = capi.EPI_OUT_F16 if eng.outer_mode == "f16" else 0      # outputs consumed by an fp16 conv
        if eng.z_rearranged:
            add(self._desc("model.0", hp, wp, self.Zbuf, self.E_lr, out_f32=self.T_fea, out_bf16=self.buf(0)))
        else:
            add(self._desc("model.0", hp, wp, self.E_fea, out_f32=self.T_fea, out_bf16=self.buf(0)))
        g = 0
        for r in range(eng.nb):
            rin = self.T_fea if r == 0 else self.R[r % 2]
            rout = self.R[(r + 1) % 2]
            for d in (1, 2, 3):
                b = self.buf(g)
                xin = rin if d == 1 else self.T[d % 2]
                pre = "model.1.sub.%d.RDB%d.convs." % (r, d)
                if self.fuse_rdb:
                    per = int(os.environ.get("ESR_RDB_LAYERS", 4))       # growth convs per fused launch (4, 2 or 1)
                    groups = [list(range(i0, min(i0 + per, 4))) for i0 in range(0, 4, per)]
                    for gi, grp in enumerate(groups):
                        self.ops.append(self._growth_desc([(eng.convs[pre + "%d.0" % i], NF + GC * i) for i in grp], b,
                                                          self.E_lat, g * len(groups) + gi, n_rdb * len(groups)))
                else:
                    for i in range(4):
                        add(self._desc(pre + "%d.0" % i, hp, wp, b, self.E_lat, flags=L, out_bf16=b, out_choff=NF + GC * i))
                last_rdb = (r == eng.nb - 1 and d == 3)
                kw = dict(flags=F16 if last_rdb else 0, alpha=0.2, beta=0.2, res2=rin if d == 3 else None,
                          out_bf16=self.buf(g + 1), out_choff=0, lo_choff=lo if last_rdb else -1)
                if self.trunk_hilo:
                    # RDB input = bf16 hi (channels 0..63 of this block's buffer) + bf16 lo, except at the RRDB input (fp32)
                    kw.update(res1=xin if d == 1 else None, res1_hilo=None if d == 1 else (b, self.T[d % 2]),
                              out_f32=rout if d == 3 else None, out_lo=None if d == 3 else self.T[(d + 1) % 2])
                else:
                    kw.update(res1=xin, out_f32=rout if d == 3 else self.T[(d + 1) % 2])
                add(self._desc(pre + "4.0", hp, wp, b, self.E_lat, **kw))
                g += 1
        trunk_out = self.buf(g)
        names = eng.outer_names
        add(self._desc(names[0], hp, wp, trunk_out, self.E_lat, flags=F16, alpha=1.0, res1=self.T_fea,
                       out_bf16=self.U[0], lo_choff=lo, up=2))
        H, W = 2 * hp, 2 * wp
        for u in range(eng.n_up):
            last = u == eng.n_up - 1
            dst = self.V1 if last else self.U[u + 1]
            add(self._desc(names[1 + u], H, W, self.U[u], flags=L | F16, out_bf16=dst, lo_choff=lo, up=1 if last else 2))
            if not last:
                H, W = 2 * H, 2 * W
        add(self._desc(names[-2], H, W, self.V1, self.E_lath, flags=L | F16, out_bf16=self.V2, lo_choff=lo))
        add(self._desc(names[-1], H, W, self.V2, self.E_lath, out_nchw=self.y))
        for d in self.ops:
            if isinstance(d, RdbGrowthDesc):
                capi.check(capi.lib().esr_seq_add_rdb_growth(self.seq, C.byref(d)))
            else:
                capi.check(capi.lib().esr_seq_add_conv(self.seq, C.byref(d), 1 if use_simt else 0))
        self.fea_x = _xslot_array(eng.fea_xslots)
        self.lat_x = _xslot_array(eng.lat_xslots)
        self.lr_x = _xslot_array(eng.lr_xslots) if eng.z_rearranged else None

    # ------------------------------------------------------------------ run
    def run_prep(self, model_input):
        """Padding / latent unpacking / row expansion of the packed [Z.view, LR] input."""
        eng, l, st = self.eng, capi.lib(), capi.stream_ptr()
        B, hp, wp, sf = self.B, self.hp, self.wp, self.sf
        capi.check(l.esr_g_input_prep(capi.ptr(model_input), B, eng.nz_in, self.h, self.w, self.m, sf,
                                      capi.ptr(self.lr_pad), capi.ptr(self.fea_in), capi.ptr(self.z_hr),
                                      capi.ptr(self.z_lr), st))
        capi.check(l.esr_expand_rows(capi.ptr(self.fea_in), B, eng.nz_in + 3, hp, wp, self.fea_x, len(self.fea_x),
                                     capi.ptr(self.E_fea), st))
        if eng.nz:
            capi.check(l.esr_expand_rows(capi.ptr(self.z_lr), B, eng.nz, hp, wp, self.lat_x, 32, capi.ptr(self.E_lat), st))
            capi.check(l.esr_expand_rows(capi.ptr(self.z_hr), B, eng.nz, sf * hp, sf * wp, self.lat_x, 32,
                                         capi.ptr(self.E_lath), st))

    def run_prep_rearranged(self, lr_pad, z):
        """first_layer_HR_rearranged: `lr_pad` [B,3,hp,wp] (already padded) and Z [B,Cz*sf^2,hp,wp] given at LR resolution."""
        eng, l, st = self.eng, capi.lib(), capi.stream_ptr()
        self.lr_pad.copy_(lr_pad)
        zb = self.Zbuf.view(torch.float16) if eng.outer_mode == "f16" else self.Zbuf       # the conv's operand format
        zb[..., :eng.z_rearranged].copy_(z.permute(0, 2, 3, 1))
        capi.check(l.esr_expand_rows(capi.ptr(self.lr_pad), self.B, 3, self.hp, self.wp, self.lr_x, 32, capi.ptr(self.E_lr), st))

    def run_convs(self):
        capi.check(capi.lib().esr_seq_run(self.seq, capi.stream_ptr()))

    def run_g(self, model_input):
        """model_input: contiguous f32 CUDA [B, 16*nz+3, h, w].  Fills self.y (raw generator output)."""
        self.run_prep(model_input)
        self.run_convs()
        return self.y

    def launches_per_forward(self, with_cem):
        eng = self.eng
        prep = 1 + 1 + (1 if eng.nz_in else 0) + (2 if eng.nz else 0)   # lr_pad, fea_in(lr [+z_lr]), z_hr, z_lr
        expand = 1 + (2 if eng.nz else 0)
        return prep + expand + capi.lib().esr_seq_num_launches(self.seq) + (2 if with_cem else 0)   # CEM x4: Down, K+Up
