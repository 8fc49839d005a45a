"""Data-gradient backward of G(+CEM): dL/d(model_input[:, :16*nz]) given dL/d(output).

The reference obtains it from autograd with every generator parameter frozen
(codes/Z_optimization.py:545-553, :633); here it is an explicit sequence of the same sm_100a kernels:
conv dgrad = the tcgen05 conv on the transposed + flipped weights, with one launch per forward conv that
produces the gradients of every dense-block slice it read plus the 9 row-expanded latent rows.

Gradient buffers (DESIGN.md "backward"):
  GF32  f32 blocked [B, 4*224/8, hp, wp, 8]: four rotating dense-block gradient frames
        [x0 64 | x1..x4 4x32 | latent 32]; RDB g uses frame g % 4 so that the RDB / RRDB residual
        gradients it needs (frames (g+1)%4 and (3r+3)%4) are still intact;
  GB    bf16 NHWC [B,hp,wp,192]: the MMA operand of the next dgrad: channels 0..63 = d(conv5 out),
        64+32(k-1).. = d(pre-activation of x_k) = LeakyReLU'(x_k) * d(x_k).
"""
import ctypes as C
import os

import torch

from . import _capi as capi
from ._capi import ConvDesc
from .engine import PackedConv, DY_ALL, NF, GC, _xslot_array

FRAME = 224          # channels of one gradient frame
LAT_OFF = 192        # latent rows inside a frame


def _lat_rows(nz):
    rows = [(c, 2 - dy) for dy in range(3) for c in range(nz)]     # (input channel, filter row pre-flipped)
    return rows + [(-1, -1)] * (32 - len(rows))


class DgradSpecs:
    """Geometry-independent packed dgrad weights of one GEngine."""

    def __init__(self, eng):
        self.eng = eng
        self.convs = {}
        nz, p = eng.nz, eng.precise_bwd
        hi_lo = ((0, 0), (64, 0), (0, 1)) if p else ((0, 0),)

        def outer_blocks():
            kb, sl = [], []
            for base, term in hi_lo:
                for c0 in (0, 32):
                    kb.append((0, base + c0, DY_ALL, 0b11))
                    sl += [(c0 + k, -1, term) for k in range(32)]
            return kb, sl

        # Trunk: the backward of an RDB is itself a dense block on the gradients.  With g_m = d(pre-activation of
        # conv m) (m = 1..4) and g_5 = d(conv5 output) stored in one bf16 buffer GB = [g5 | g1 | g2 | g3 | g4],
        #   d(x_j) = sum_{m > j} W_m[:, x_j]^T (*) g_m          (conv m reads x_0 .. x_{m-1}: block.py:230-235)
        # is ONE conv launch per j whose K runs over the g_m it needs (accumulated in TMEM, no fp32 read-modify-write
        # fan-in) and whose N is the slice x_j.  Launch "k" (k = 5..1) produces d(x_{k-1}); k = 1 also emits the 9
        # row-expanded latent rows of the whole block.  Weights: U[co_slot, ci] = W_m[co, ci] gathered per RDB.
        self.trunk = {}
        for r in range(eng.nb):
            for d in (1, 2, 3):
                pre = "model.1.sub.%d.RDB%d.convs." % (r, d)
                for k in (5, 4, 3, 2, 1):
                    ms = [5] + [m for m in (4, 3, 2, 1) if m >= k]          # convs that read x_{k-1}
                    kb, sl = [], []
                    for m in ms:
                        base = 0 if m == 5 else NF + GC * (m - 1)           # channels of g_m inside GB
                        for c0 in range(0, NF if m == 5 else GC, 32):
                            kb.append((0, base + c0, DY_ALL, 0b11))
                            sl += [(base + c0 + c, -1, 0) for c in range(32)]
                    if k >= 2:
                        rows = [(nz + NF + GC * (k - 2) + c, -1) for c in range(GC)]
                    else:
                        rows = [(nz + c, -1) for c in range(NF)] + (_lat_rows(nz) if nz else [])
                    self.trunk[pre + "%d" % k] = PackedConv(pre + "dgrad%d" % k, len(rows), kb, sl, rows, 32, pair=eng.pair)
        names = eng.outer_names
        for name in names[:-1]:
            is_up = name in eng.upconv_names
            has_lat = nz > 0 and not is_up
            kb, sl = outer_blocks()
            rows = [((0 if is_up else nz) + c, -1) for c in range(NF)] + (_lat_rows(nz) if has_lat else [])
            self.convs[name] = PackedConv(name + ".dgrad", len(rows), kb, sl, rows, 32, pair=eng.pair)
        # last conv (out_nc -> 64 [+latent]): its input gradient arrives as NCHW f32 and is row-"expanded"
        # (dy = 0 only) into [hi | lo | hi] slots
        onc = eng.out_nc
        self.gy_xslots = [(c, 0, t) for t in (0, 1, 0) for c in range(onc)]
        self.gy_xslots += [(-1, 0, 0)] * (32 - len(self.gy_xslots))
        sl = [(c, -1, t) for t in (0, 0, 1) for c in range(onc)] + [(-1, -1, 0)] * (32 - 3 * onc)
        rows = [(nz + c, -1) for c in range(NF)] + (_lat_rows(nz) if nz else [])
        self.convs[names[-1]] = PackedConv(names[-1] + ".dgrad", len(rows), [(0, 0, DY_ALL, 0b01 if 3 * onc <= 16 else 0b11)],
                                           sl, rows, 32, pair=eng.pair)
        # first conv: only the latent rows are needed (the LR image takes no gradient)
        kb, sl = outer_blocks()
        self.convs["model.0"] = PackedConv("model.0.dgrad", 32, kb, sl, _lat_rows(eng.nz_in), 32, pair=eng.pair) if eng.nz_in else None

    def pack(self, params):
        """Packs every dgrad weight image: one segment-copy launch gathers each dense block's five weight tensors into its
        [co_slot, ci, 3, 3] array U, one table launch packs all ~420 images (built once per set of source pointers)."""
        from .engine import PackTable
        from ._capi import CopySeg
        eng = self.eng
        key = tuple(params[name][0].data_ptr() for name in sorted(params))
        tab = getattr(self, "_pack_table", None)
        if tab is None or tab.key != key:
            dev = next(iter(params.values()))[0].device
            entries, segs = [], []
            self._U = {}
            for name, pc in self.convs.items():
                if pc is None:
                    continue
                w, _ = params[name]
                cin = w.shape[1]
                # logical dgrad weight [row = ci, slot = co, ky, kx] = W[co, ci, 2-ky, 2-kx]
                pc.pack(w, None, 8, 9, cin * 9, -3, -1, table=entries)
            for r in range(eng.nb):
                for d in (1, 2, 3):
                    pre = "model.1.sub.%d.RDB%d.convs." % (r, d)
                    ws = [params[pre + "%d.0" % i][0] for i in range(5)]
                    ci = ws[4].shape[1]
                    U = self._U[pre] = torch.zeros(NF + 4 * GC, ci, 3, 3, dtype=torch.float32, device=dev)   # [co_slot, ci, ky, kx]
                    for m in range(5):          # U[:64] = W5, U[64 + 32(m-1) ..][:, :cin_m] = W_m
                        w = ws[4] if m == 0 else ws[m - 1]
                        row0 = 0 if m == 0 else NF + GC * (m - 1)
                        sg = CopySeg()
                        sg.src, sg.dst = w.data_ptr(), U.data_ptr() + 4 * row0 * ci * 9
                        sg.rows, sg.row_elems, sg.src_pitch, sg.dst_pitch = w.shape[0], w.shape[1] * 9, w.shape[1] * 9, ci * 9
                        segs.append(sg)
                    for k in (5, 4, 3, 2, 1):
                        self.trunk[pre + "%d" % k].pack(U, None, 8, 9, ci * 9, -3, -1, table=entries)
            arr = (CopySeg * len(segs))(*segs)
            self._segs = (torch.frombuffer(bytearray(bytes(memoryview(arr))), dtype=torch.uint8).to(dev), len(segs))
            tab = self._pack_table = PackTable()
            tab.build(key, entries, dev)
        capi.check(capi.lib().esr_copy_segments(capi.ptr(self._segs[0]), self._segs[1], capi.stream_ptr()))
        tab.run()


class BackwardPlan:
    """Buffers and recorded dgrad sequences for one GPlan (built with keep_activations=True)."""

    def __init__(self, plan, specs, use_simt=False, keep_gb=False):
        self.plan, self.specs = plan, specs
        self.keep_gb = keep_gb          # training: every RDB keeps its own gradient buffer (the weight gradients read them
                                        # after the data-gradient pass); Z optimisation: two buffers take turns
        eng = plan.eng
        B, hp, wp, sf, dev = plan.B, plan.hp, plan.wp, plan.sf, plan.device
        f32 = dict(dtype=torch.float32, device=dev)
        bf = dict(dtype=torch.bfloat16, device=dev)
        self.GF32 = torch.zeros(B * 4 * FRAME * hp * wp, **f32)
        self.GBs = [torch.zeros(B, hp, wp, 192, **bf) for _ in range(3 * eng.nb if keep_gb else 2)]   # RDB g works in gb(g); its
                                                                         # d(x_0) launch reads all of it and writes the next g_5 into gb(g-1)
        self.GS = torch.zeros(B, hp, wp, 128, **bf)
        self.Gsc = torch.zeros(B * 64 * hp * wp, **f32)
        self.GFea = torch.zeros(B, hp, wp, 128, **bf)
        H4, W4 = sf * hp, sf * wp
        self.g_y = torch.zeros(B, eng.out_nc, H4, W4, **f32)
        self.E6 = torch.zeros(B, H4, W4, 32, **bf)
        self.GH = torch.zeros(B * 96 * H4 * W4, **f32)
        self.GV = torch.zeros(B, H4, W4, 128, **bf)
        self.GV1 = torch.zeros(B, H4, W4, 128, **bf)
        self.GU, self.GVu = [], []          # per upconv: f32 dgrad output at its resolution, bf16 input of the one below
        for u in range(eng.n_up):
            res = 2 ** (u + 1)
            self.GU.append(torch.zeros(B * 64 * res * hp * res * wp, **f32))
            self.GVu.append(torch.zeros(B, res * hp, res * wp, 128, **bf) if u < eng.n_up - 1 else self.GV1)
        self.g_z_lr = torch.zeros(B, max(eng.nz_in, 1), hp, wp, **f32)
        self.g_z_hr = torch.zeros(B, max(eng.nz, 1), H4, W4, **f32)
        self.gy_x = _xslot_array(specs.gy_xslots)
        self.lat_x = _xslot_array([(c, dy - 1, 0) for dy in range(3) for c in range(max(eng.nz_in, 1))])
        # backward twin of the fused growth launch (engine.GPlan.fuse_rdb): the four Mask dgrads of an RDB as one persistent
        # launch with tile-level dependencies (mode 1 of esr_rdb_growth_tc); its own dependency counters
        self.fuse_rdb = bool(getattr(plan, "fuse_rdb", False)) and not use_simt and os.environ.get("ESR_FUSE_RDB_BWD", "1") != "0"
        self.rdb_flags = torch.zeros(int(capi.lib().esr_rdb_growth_flag_words(B, hp, wp)), dtype=torch.int32, device=dev) \
            if self.fuse_rdb else None
        self.steps = []          # ("seq", handle) | ("call", fn)
        self._seqs = []
        self._cur = None
        self._descs = []
        self.use_simt = use_simt
        self._record()

    def __del__(self):
        try:
            for s in self._seqs:
                capi.lib().esr_seq_destroy(s)
        except Exception:
            pass

    def gb(self, g):
        return self.GBs[g] if self.keep_gb else self.GBs[g % 2]

    # ------------------------------------------------------------------ recording helpers
    def _conv(self, name, H, W, src, out_f32, out_stride, out_choff=0, accum=False, no_accum=0, lat_tile_to=None,
              out_bf16=None, bf16_stride=0, lo_choff=-1, scale=1.0, only_bf16_tiles=None, mask=None, mask_stride=0,
              res1=None, res1_choff=0, gamma=1.0, res2=None, res2_choff=0, no_res=0, pc=None, bf16_choff=0, mask_choff=0):
        pc = pc if pc is not None else self.specs.convs[name]
        d = ConvDesc()
        d.B, d.H, d.W = self.plan.B, H, W
        d.src[0].ptr, d.src[0].channels = src.data_ptr(), src.shape[-1]
        d.cout_tile, d.cout_tiles, d.num_kblocks, d.pair = pc.cout_tile, pc.cout_tiles, pc.nkb, pc.pair
        for i in range(pc.nkb):
            d.kblocks[i] = pc.kblocks[i]
        d.wpack, d.w_tile_bytes, d.bias = pc.wpack.data_ptr(), pc.w_tile_bytes, pc.bias.data_ptr()
        d.flags = capi.EPI_F32_BLOCKED | (capi.EPI_ACCUM if accum else 0)
        d.no_accum_tiles = no_accum
        if out_f32 is not None:
            d.out_f32, d.out_f32_stride, d.out_f32_choff = out_f32.data_ptr(), out_stride, out_choff
        if lat_tile_to is not None:
            d.tile_choff[pc.cout_tiles - 1] = lat_tile_to
        if out_bf16 is not None:
            d.out_bf16, d.out_bf16_stride, d.out_bf16_choff = out_bf16.data_ptr(), bf16_stride, bf16_choff
            d.out_bf16_lo_choff, d.out_bf16_scale = lo_choff, scale
            if only_bf16_tiles is not None:
                allt = (1 << pc.cout_tiles) - 1
                keep = 0
                for t in only_bf16_tiles:
                    keep |= 1 << t
                d.no_bf16_tiles = allt & ~keep
        if mask is not None:
            d.flags |= capi.EPI_MASK
            d.mask, d.mask_stride, d.mask_choff = mask.data_ptr(), mask_stride, mask_choff
        if res1 is not None:
            d.flags |= capi.EPI_RES1
            d.res1, d.res1_stride, d.res1_choff, d.gamma = res1.data_ptr(), out_stride, res1_choff, gamma
        if res2 is not None:
            d.flags |= capi.EPI_RES2
            d.res2, d.res2_stride, d.res2_choff, d.beta = res2.data_ptr(), out_stride, res2_choff, 1.0
        d.no_res_tiles = no_res
        self._descs.append(d)
        if os.environ.get("ESR_BWD_EAGER") == "1":         # debug: one launch + sync per conv, names the failing one
            tag = (name or pc.name, H, W)
            self.steps.append(("conv", (d, tag)))
            return
        if self._cur is None:
            self._cur = capi.lib().esr_seq_create()
            self._seqs.append(self._cur)
            self.steps.append(("seq", self._cur))
        capi.check(capi.lib().esr_seq_add_conv(self._cur, C.byref(d), 1 if self.use_simt else 0))

    def _call(self, fn):
        self._cur = None
        self.steps.append(("call", fn))

    def _combine(self, src, src_stride, src_choff, pool, add, H, W, out_f32, mask, mask_stride, mask_sub, scale, out_bf16, lo):
        l, B = capi.lib(), self.plan.B

        def run():
            capi.check(l.esr_grad_combine(capi.ptr(src), src_stride, src_choff, pool, capi.ptr(add), 64, 0, B, H, W,
                                          capi.ptr(out_f32), 64, 0, capi.ptr(mask), mask_stride, 0, mask_sub, 0.2, scale,
                                          capi.ptr(out_bf16), 128, 0, 64 if lo else -1, capi.stream_ptr()))
        self._call(run)

    # ------------------------------------------------------------------ the backward graph
    def _record(self):
        plan, eng = self.plan, self.plan.eng
        hp, wp, sf = plan.hp, plan.wp, plan.sf
        H4, W4 = sf * hp, sf * wp
        names = eng.outer_names
        lo = 64 if eng.precise_bwd else -1
        nz = eng.nz
        lat_bit = lambda pc: 1 << (pc.cout_tiles - 1)
        # ---- HR convs
        pc = self.specs.convs[names[-1]]
        self._conv(names[-1], H4, W4, self.E6, self.GH, 96, out_bf16=self.GV, bf16_stride=128, lo_choff=lo,
                   only_bf16_tiles=(0, 1), mask=plan.V2, mask_stride=plan.V2.shape[-1])
        self._conv(names[-2], H4, W4, self.GV, self.GH, 96, accum=nz > 0, no_accum=0b011, out_bf16=self.GV1,
                   bf16_stride=128, lo_choff=lo, only_bf16_tiles=(0, 1), mask=plan.V1, mask_stride=plan.V1.shape[-1])
        # ---- upconvs, top down; each followed by the 2x2 sum-pool adjoint of the nearest upsample
        for u in reversed(range(eng.n_up)):
            res = 2 ** (u + 1)
            self._conv(names[1 + u], res * hp, res * wp, self.GVu[u], self.GU[u], 64)
            lowH, lowW = res // 2 * hp, res // 2 * wp
            if u > 0:      # below sits upconv u-1, whose LeakyReLU output was stored 2x2-replicated in plan.U[u]
                self._combine(self.GU[u], 64, 0, 2, None, lowH, lowW, None, plan.U[u], plan.U[u].shape[-1], 2, 1.0, self.GVu[u - 1], eng.precise_bwd)
            else:          # below sits the trunk shortcut sum (no activation)
                self._combine(self.GU[0], 64, 0, 2, None, lowH, lowW, self.Gsc, None, 0, 1, 1.0, self.GS, eng.precise_bwd)
        # ---- LR_conv: H = d(last RRDB output) -> frame (3nb)%4, emits 0.04*H for RDB3.conv5 of the last RRDB
        n_rdb = 3 * eng.nb
        touched = set()
        q0 = n_rdb % 4
        touched.add(q0)
        self._conv(names[0], hp, wp, self.GS, self.GF32, 4 * FRAME, out_choff=FRAME * q0, lat_tile_to=LAT_OFF if nz else None,
                   out_bf16=self.gb(n_rdb - 1), bf16_stride=192, scale=0.04, only_bf16_tiles=(0, 1))
        # ---- the trunk, last RDB first
        for g in reversed(range(n_rdb)):
            r, dd = divmod(g, 3)            # dd = 0,1,2 -> RDB1,2,3
            q = g % 4
            first_touch = q not in touched
            touched.add(q)
            buf = plan.bufs[g]
            GB = self.gb(g)
            pre = "model.1.sub.%d.RDB%d.convs." % (r, dd + 1)
            if self.fuse_rdb and os.environ.get("ESR_BWD_EAGER") != "1":
                # the four launches below as ONE persistent launch: layer k may read what layers > k wrote once the 3x3 tile
                # neighbourhood is stored (same kernel as the forward's fused growth convs, Mask epilogue)
                layers = [(self.specs.trunk[pre + "%d" % k], NF + GC * (k - 2)) for k in (5, 4, 3, 2)]
                d = plan._growth_desc(layers, GB, None, n_rdb - 1 - g, n_rdb, mode=1, mask=buf, flags=self.rdb_flags)
                if self._cur is None:
                    self._cur = capi.lib().esr_seq_create()
                    self._seqs.append(self._cur)
                    self.steps.append(("seq", self._cur))
                self._descs.append(d)
                capi.check(capi.lib().esr_seq_add_rdb_growth(self._cur, C.byref(d)))
            else:
                for k in (5, 4, 3, 2):
                    # d(x_{k-1}) complete in one launch (K over g_5, g_4 .. g_k); emit g_{k-1} = LeakyReLU'(x_{k-1}) * d(x_{k-1})
                    ch = NF + GC * (k - 2)
                    self._conv(None, hp, wp, GB, None, 0, pc=self.specs.trunk[pre + "%d" % k], out_bf16=GB, bf16_stride=192,
                               bf16_choff=ch, mask=buf, mask_stride=192, mask_choff=ch)
            # d(x_0): all five convs read it; add the residual paths, emit 0.2x (0.04x across an RRDB boundary) as the
            # g_5 of the RDB below; the last cout tile holds the block's 9 latent rows (accumulated per frame)
            pc = self.specs.trunk[pre + "1"]
            lat = lat_bit(pc) if nz else 0
            main_bits = ((1 << pc.cout_tiles) - 1) & ~lat
            kw = dict(out_choff=FRAME * q, lat_tile_to=LAT_OFF if nz else None, accum=nz > 0,
                      no_accum=main_bits | (lat if first_touch else 0),
                      res1=self.GF32, res1_choff=FRAME * ((g + 1) % 4), gamma=0.2 if dd == 2 else 1.0, no_res=lat)
            if dd == 0:
                kw.update(res2=self.GF32, res2_choff=FRAME * ((3 * r + 3) % 4))
            if g > 0:
                kw.update(out_bf16=self.gb(g - 1), bf16_stride=192, only_bf16_tiles=(0, 1), scale=0.04 if dd == 0 else 0.2)
            self._conv(None, hp, wp, GB, self.GF32, 4 * FRAME, pc=pc, **kw)
        # ---- d(fea) = d(RRDB0 input) + shortcut gradient -> first conv's latent rows
        if eng.nz_in:
            self._combine(self.GF32, 4 * FRAME, 0, 1, self.Gsc, hp, wp, None, None, 0, 1, 1.0, self.GFea, eng.precise_bwd)
            self._conv("model.0", hp, wp, self.GFea, self.GF32, 4 * FRAME, out_choff=0, lat_tile_to=LAT_OFF,
                       accum=nz > 0)
        self.n_lat_acc = 4 if nz else 1

    # ------------------------------------------------------------------ run
    def run(self, g_y):
        """g_y: f32 [B,out_nc,4hp,4wp] gradient w.r.t. the raw generator output.  Returns d(model_input)."""
        plan, eng, l, st = self.plan, self.plan.eng, capi.lib(), capi.stream_ptr()
        B, hp, wp, sf = plan.B, plan.hp, plan.wp, plan.sf
        H4, W4 = sf * hp, sf * wp
        capi.check(l.esr_expand_rows(capi.ptr(g_y), B, eng.out_nc, H4, W4, self.gy_x, 32, capi.ptr(self.E6), st))
        for kind, obj in self.steps:
            if kind == "seq":
                capi.check(l.esr_seq_run(obj, capi.stream_ptr()))
            elif kind == "conv":
                try:
                    capi.check(l.esr_conv3x3_tc(C.byref(obj[0]), capi.stream_ptr()))
                    torch.cuda.synchronize()
                except Exception as e:
                    raise capi.EsrError("dgrad launch %s failed: %s" % (obj[1], e))
            else:
                obj()
        g_in = torch.empty(B, eng.nz_in * sf * sf + 3, plan.h, plan.w, dtype=torch.float32, device=plan.device)
        nlat = 3 * eng.nz_in
        capi.check(l.esr_expand_rows_bwd(capi.ptr(self.GF32), 4 * FRAME, LAT_OFF, 1, self.n_lat_acc, FRAME, B, eng.nz_in, hp, wp,
                                         self.lat_x, nlat, capi.ptr(self.g_z_lr), st))
        g_hr = None
        if eng.nz:
            capi.check(l.esr_expand_rows_bwd(capi.ptr(self.GH), 96, 64, 1, 1, 0, B, eng.nz, H4, W4, self.lat_x, nlat,
                                             capi.ptr(self.g_z_hr), st))
            g_hr = self.g_z_hr
        capi.check(l.esr_g_input_prep_bwd(capi.ptr(g_hr), capi.ptr(self.g_z_lr), B, eng.nz_in, plan.h, plan.w, plan.m, sf,
                                          capi.ptr(g_in), st))
        return g_in

    def num_launches(self):
        n = 1 + 3
        for kind, obj in self.steps:
            n += capi.lib().esr_seq_num_launches(obj) if kind == "seq" else 1
        return n


def generator_backward_eager(plan, bp, cem_filters, margin, g):
    """d(model_input) for the gradient g w.r.t. the (cropped) output of G+CEM; g: contiguous f32 on the plan's device."""
    eng = plan.eng
    B, sf = plan.B, plan.sf
    H4, W4 = sf * plan.hp, sf * plan.wp
    if cem_filters is not None:
        n = B * eng.out_nc * (H4 * W4 + H4 * plan.wp + 2 * plan.hp * plan.wp)
        ws = torch.empty(n, dtype=torch.float32, device=g.device)
        capi.cem_call("project_bwd", cem_filters, capi.ptr(g), B, eng.out_nc, H4, W4, sf * margin,
                      capi.ptr(bp.g_y), capi.ptr(ws), capi.stream_ptr())
        g_y = bp.g_y
    else:
        g_y = g
    return bp.run(g_y)


def generator_backward(ctx, g):
    """autograd hook of rrdbnet._GeneratorFn."""
    plan, net = ctx.plan, ctx.net
    if len(plan.bufs) <= 2:
        raise capi.EsrError("backward needs a forward pass recorded with gradients enabled")
    if plan.eng.nz_in == 0:
        raise capi.EsrError("this generator has no latent input: nothing to differentiate")
    bp = net.backward_plan(plan)
    with torch.cuda.device(g.device):
        return generator_backward_eager(plan, bp, ctx.cem_filters, ctx.margin, g.contiguous().float())
