"""The GAN training step, data parallel (BASELINE config 5): ``GeneratorTrainer`` (the generator's
half: forward with kept activations, data + weight gradients, bucketed NCCL all-reduce) and ``GanTrainer`` (the whole
iteration with the critic of esr_b200.discriminator and the WGAN-GP penalty).

The reference's ``SRRaGANModel.optimize_parameters`` (codes/models/SRRaGAN_model.py:463-547) runs
``fake_H = netG(model_input)``, builds ``l_g_total`` from ``fake_H`` (pixel / feature / GAN / range terms) and calls
``l_g_total.backward(); optimizer_G.step()`` with every generator parameter trainable, under ``nn.DataParallel``
(models/networks.py:99-101).  Here the generator's part of that is explicit:

    trainer = GeneratorTrainer(netG)                  # netG: this package's CEM_PyTorch(RRDBNet), parameters on the GPU
    fake_H = trainer.forward(model_input)             # leaf tensor: build any torch loss on it (the reference's own
    loss(fake_H).backward()                           #   criteria, a discriminator, ...) and back-propagate to fake_H
    trainer.backward(fake_H.grad)                     # dL/dW, dL/db of all 351 convs into p.grad, averaged over the ranks
    optimizer_G.step()

``backward`` = the data-gradient pass of Z optimisation (backward.py: the same tcgen05 dgrad launches, with one gradient
buffer kept per dense block) followed by the weight-gradient kernels of csrc/wgrad.cu over the activations / gradients
both passes left in HBM.  One process per GPU: all gradients live in ONE flat fp32 buffer ordered last layer first; it is
cut into buckets, and as soon as a bucket's weight-gradient launches are queued its ``all_reduce`` (NCCL, average) is
issued on a communication stream, so the exchange of the late layers runs under the weight-gradient kernels of the
early ones (68.2 MB for the production generator).  ``p.grad`` of every parameter is a view into that buffer.

The critic, the WGAN-GP double backward and the range / pixel criteria of the full step are torch code on ``fake_H``
(``GanTrainer`` below, or the reference's own ``optimize_parameters`` through the autograd node rrdbnet._TrainFn).
Not built: the VGG feature loss, and a tcgen05 form of the weight-gradient GEMM (csrc/wgrad.cu uses warp-level bf16 MMAs).
"""
import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _capi as capi
from ._capi import WgradItem, WgradSmallItem, WgradTcItem
from .backward import BackwardPlan, DgradSpecs, generator_backward_eager
from .cem import CEM_PyTorch
from .engine import NF, GC, _struct_array_to_device
from .rrdbnet import RRDBNet, _forward_eager

TILE_H, TILE_W = 8, 16


class GeneratorTrainer:
    def __init__(self, netG, bucket_bytes=8 << 20, process_group=None, attach_grads=True):
        wrapper = getattr(netG, 'module', netG)
        G = getattr(wrapper, 'generated_image_model', wrapper)
        if not isinstance(G, RRDBNet):
            raise capi.EsrError("GeneratorTrainer needs this package's RRDBNet (optionally wrapped by CEM_PyTorch)")
        if G._cfg['nz_in'] == 0:
            raise NotImplementedError("training a generator without an HR_downscaled / LR latent input is not built (the "
                                      "production configuration has one)")
        self.wrapper = wrapper if isinstance(wrapper, CEM_PyTorch) else None
        self.G, self.group, self.bucket_bytes = G, process_group, bucket_bytes
        self.dev = next(G.parameters()).device
        if self.dev.type != 'cuda':
            raise capi.EsrError("GeneratorTrainer: move the generator to a B200 first (no CPU path)")
        self.comm = torch.cuda.Stream(device=self.dev)
        # parameters in backward order (last conv first): the flat gradient buffer and its buckets follow it
        named = dict(G.named_parameters())
        eng = G.engine()
        order = list(reversed(eng.outer_names[1:])) + [eng.outer_names[0]]
        for r in reversed(range(eng.nb)):
            for d in (3, 2, 1):
                order += ["model.1.sub.%d.RDB%d.convs.%d.0" % (r, d, i) for i in (4, 3, 2, 1, 0)]
        order.append("model.0")
        assert sorted(order) == sorted(eng.convs.keys())
        self.order = order
        total, self.slices = 0, {}
        for name in order:
            for suffix in (".weight", ".bias"):
                p = named[name + suffix]
                self.slices[name + suffix] = (total, p.numel(), p)
                total += (p.numel() + 3) & ~3
        self.flat = torch.zeros(total, dtype=torch.float32, device=self.dev)
        if attach_grads:                 # explicit API: p.grad IS the flat buffer.  The autograd node (rrdbnet._TrainFn)
            for key, (off, n, p) in self.slices.items():   # hands copies to autograd's own accumulation instead
                p.grad = self.flat[off:off + n].view_as(p)
        # buckets: contiguous ranges of `order`
        self.buckets, cur, cur_bytes, start = [], [], 0, 0
        for name in order:
            cur.append(name)
            cur_bytes += 4 * (self.slices[name + ".weight"][1] + self.slices[name + ".bias"][1])
            if cur_bytes >= bucket_bytes:
                end = self.slices[name + ".bias"][0] + ((self.slices[name + ".bias"][1] + 3) & ~3)
                self.buckets.append((cur, start, end))
                cur, cur_bytes, start = [], 0, end
        if cur:
            self.buckets.append((cur, start, total))
        self._state = None
        self._tables = {}
        self.split_k = os.environ.get("ESR_WGRAD_SPLITK", "1") != "0"      # esr_wgrad16r (0: the fragment-per-warp kernel, A/B timing)
        # esr_wgrad_tc: the dense-block convs' 16-bit input channels on the tcgen05 path (0: everything on the mma.sync kernels)
        self.use_tc = os.environ.get("ESR_WGRAD_TC", "1") != "0"
        self.tc_chunks = int(os.environ.get("ESR_WGRAD_TC_CHUNKS", "2"))
        # one launch per kernel for the whole network (two halves when gradients are exchanged, so that the first half's
        # all-reduce runs under the second half's kernels) instead of one per gradient bucket: a bucket's 50-80 CTAs leave half
        # of the 148 SMs idle, which costs more than the finer overlap gains
        self.per_bucket = os.environ.get("ESR_WGRAD_PER_BUCKET", "0") == "1"

    # ------------------------------------------------------------------ forward
    def forward(self, model_input, margin=None, filters=None, leaf=True):
        """fake_H = netG(model_input) with the activations kept; returns a leaf tensor that requires grad.
        margin / filters: given by the autograd node, which is called below the CEM wrapper (leaf=False there)."""
        G, w = self.G, self.wrapper
        x = model_input.contiguous().float()
        capi.require_device(self.dev.index if self.dev.index is not None else torch.cuda.current_device())
        B, _, h, wd = x.shape
        if margin is None:
            margin = (w._margin_LR if w.pre_pad else 0) if w is not None else 0
            filters = w._filters if w is not None else None
        with torch.cuda.device(self.dev):
            G.weights_changed()                              # training: the optimiser wrote the weights since the last forward
            plan = G.plan(B, h, wd, margin, keep=True, slot='train')
            sf = G.upscale
            crop = sf * margin
            onc = plan.y.size(1)
            out = torch.empty(B, onc, sf * plan.hp - 2 * crop, sf * plan.wp - 2 * crop, device=self.dev, dtype=torch.float32)
            ws = torch.empty(max(1, 2 * B * onc * plan.hp * plan.wp), device=self.dev, dtype=torch.float32) if filters is not None else None
            _forward_eager(plan, x, filters, crop, out, ws)
        self._state = (plan, filters, margin)
        self.stamp = getattr(self, 'stamp', 0) + 1
        return out.requires_grad_(True) if leaf else out

    # ------------------------------------------------------------------ weight-gradient work items
    def _backward_plan(self, plan):
        G = self.G
        key = id(plan)
        ent = getattr(self, '_bp', None)
        if ent is None or ent[0] != key or ent[2] != G._engine_ptr_key:
            if G._dgrad is None:
                G._dgrad = DgradSpecs(G._engine)
                G._dgrad.pack(G._packed_params)
            self._bp = (key, BackwardPlan(plan, G._dgrad, keep_gb=True), G._engine_ptr_key)
            self._tables = {}
        return self._bp[1]

    def _items(self, plan, bp):
        """Device tables of esr_wgrad_item / esr_wgrad_small_item per bucket (built once per plan)."""
        key = (id(plan), id(bp))
        if key in self._tables:
            return self._tables[key]
        eng = plan.eng
        B, hp, wp, sf = plan.B, plan.hp, plan.wp, plan.sf
        nz, nzi = eng.nz, eng.nz_in
        f16 = 1 if eng.outer_mode == "f16" else 0
        n_rdb = 3 * eng.nb
        names = eng.outer_names
        H4, W4 = sf * hp, sf * wp
        # conv -> (x16 tensor, channels, is fp16, row-expanded small-channel tensor, its fp16 flag, small channels,
        #          g tensor, g first channel, cout staged, n_co, H, W)
        # The <= 6 small input channels (latent, LR image) are taken from the forward's row-expanded 16-bit tensors
        # (E_lat / E_lath / E_fea: slot dy * n + c holds channel c of row y + dy - 1, so the centre row dy = 1 is the
        # plain channel): one more 16-channel block on the tensor path instead of a scalar kernel over fp32 planes.
        spec = {}
        lat, lath = (plan.E_lat, 0, nz), (plan.E_lath, 0, nz)
        spec[names[-1]] = (plan.V2, NF, f16, lath, bp.E6, 0, 16, eng.out_nc, H4, W4)
        spec[names[-2]] = (plan.V1, NF, f16, lath, bp.GV, 0, NF, NF, H4, W4)
        for u in range(eng.n_up):
            res = 2 ** (u + 1)
            spec[names[1 + u]] = (plan.U[u], NF, f16, (None, 0, 0), bp.GVu[u], 0, NF, NF, res * hp, res * wp)
        spec[names[0]] = (plan.buf(n_rdb), NF, f16, lat, bp.GS, 0, NF, NF, hp, wp)
        for g in range(n_rdb):
            r, dd = divmod(g, 3)
            for i in range(5):
                cout = GC if i < 4 else NF
                gc0 = 0 if i == 4 else NF + GC * i
                spec["model.1.sub.%d.RDB%d.convs.%d.0" % (r, dd + 1, i)] = (plan.bufs[g], NF + GC * i, 0, lat, bp.gb(g), gc0,
                                                                             cout, cout, hp, wp)
        spec["model.0"] = (None, 0, 0, (plan.E_fea, f16, nzi + 3), bp.GFea, 0, NF, NF, hp, wp)
        # tensor-map table of the tcgen05 path: one map per dense-block buffer (kind 0) and per gradient buffer (kind 1)
        l = capi.lib()
        map_bytes = int(l.esr_wgrad_tc_map_bytes())
        maps_raw, map_index = [], {}

        def map_of(t, kind):
            k = (t.data_ptr(), kind)
            if k not in map_index:
                buf = (C.c_uint8 * map_bytes)()
                capi.check(l.esr_wgrad_tc_make_map(buf, C.c_void_p(t.data_ptr()), t.shape[-1], t.shape[0], t.shape[1], t.shape[2], kind))
                map_index[k] = len(maps_raw)
                maps_raw.append(bytes(buf))
            return map_index[k]
        tc_ok = self.use_tc and nz > 0
        self._bf16_copies = []                             # (fp16 view of an outer conv's input, its bf16 copy for esr_wgrad_tc)
        tables = []
        # launch groups: one per bucket (per_bucket), or the whole network at once on a single GPU, or two halves when the
        # gradients are exchanged: the first half's all-reduce then runs under the second half's kernels
        nb_ = len(self.buckets)
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
        cuts = list(range(nb_ + 1)) if self.per_bucket else ([0, (nb_ + 1) // 2, nb_] if (multi and nb_ > 1) else [0, nb_])
        self._group_cuts = cuts
        groups = [[n for b in self.buckets[a:b_] for n in b[0]] for a, b_ in zip(cuts[:-1], cuts[1:])]
        for names_b in groups:
            big, small, tcs = [], [], []
            for name in names_b:
                x16, c16, xf16, (xs, xsf16, n_c), g, gc0, cout, n_co, H, W = spec[name]
                w_off, _, wp_ = self.slices[name + ".weight"]
                b_off = self.slices[name + ".bias"][0]
                cin_total = wp_.shape[1]
                dw = self.flat.data_ptr() + 4 * w_off
                if xs is None:
                    n_c = 0
                assert cin_total == c16 + n_c, (name, cin_total, c16, n_c)
                tiles = B * ((H + TILE_H - 1) // TILE_H) * ((W + TILE_W - 1) // TILE_W)
                chunks = max(1, (H * W) // (hp * wp))          # higher-resolution convs are cut into chunks of LR-conv size
                blocks = [(x16, c0, xf16, 0, min(16, c16 - c0), n_c + c0) for c0 in range(0, c16, 16)]
                if tc_ok and x16 is not None and c16 >= 64 and cout >= 32:
                    # the 16-bit input channels go to esr_wgrad_tc in blocks of 128 x 32 output channels; what stays on the
                    # mma.sync kernel is the block that carries the bias sum: the latent block below, or - for the upconvs,
                    # which take no latent - the first 16 channels.  The outer convs' activations are fp16 and tcgen05
                    # kind::f16 rejects an fp16 x bf16 operand pair (illegal instruction, tried), while their gradients
                    # (~1e-6) do not fit fp16: they are read from a bf16 copy made at the start of backward - the rounding
                    # the mma.sync kernel applies to them while staging
                    x_tc = x16
                    if xf16:
                        x_tc = torch.empty_like(x16)                                   # bf16 storage
                        self._bf16_copies.append((x16.view(torch.float16), x_tc))
                    tc_c0 = 0 if n_c else 16
                    blocks = blocks[:tc_c0 // 16]
                    per_item = int(os.environ.get("ESR_WGRAD_TC_TILES", 128))       # tiles per CTA of a high-resolution conv
                    nchunk = max(self.tc_chunks, tiles // per_item) if tiles > 128 else self.tc_chunks
                    for c0 in range(tc_c0, c16, 128):
                        for co0 in range(0, cout, 32):
                            for ch in range(nchunk):           # tile ranges: more, shorter CTAs fill the last wave better
                                ti = WgradTcItem()
                                ti.x_map, ti.g_map = map_of(x_tc, 0), map_of(g, 1)
                                ti.x_c0, ti.g_c0, ti.x_f16 = c0, gc0 + co0, 0
                                ti.dw = dw + 4 * co0 * cin_total * 9
                                ti.n_ci, ti.n_co = min(128, c16 - c0), min(32, n_co - co0)
                                ti.cin_total, ti.ci0 = cin_total, n_c + c0
                                ti.B, ti.H, ti.W = B, H, W
                                if nchunk > 1:                 # order-independent only for two contributions; more: fp32 sum order
                                    ti.tile_begin, ti.tile_end = tiles * ch // nchunk, tiles * (ch + 1) // nchunk
                                tcs.append(ti)
                if n_c:                                        # the centre-row slots [n_c, 2 n_c) of the row-expanded tensor
                    assert 2 * n_c <= 16
                    blocks.append((xs, 0, xsf16, n_c, 2 * n_c, 0))
                last = name == names[-1]                       # its bias gradient is summed from the fp32 planes (below)
                # esr_wgrad16r keeps all nine taps' accumulators per warp: at most 32 staged output channels per item, so
                # the 64-channel convs are cut into two halves of the output channels
                halves = [(0, cout, n_co)] if (cout <= 32 or not self.split_k) else [(0, 32, 32), (32, 32, n_co - 32)]
                for bi, (src, c0, sf16, ci_lo, ci_hi, ci0) in enumerate(blocks):
                    for ch in range(chunks):
                        for (co0, co_staged, co_valid) in halves:
                            it = WgradItem()
                            it.x, it.g, it.dw = src.data_ptr(), g.data_ptr(), dw + 4 * co0 * cin_total * 9
                            it.x_stride, it.x_c0, it.x_f16 = src.shape[-1], c0, sf16
                            it.g_stride, it.g_c0, it.cout = g.shape[-1], gc0 + co0, co_staged
                            it.n_co, it.n_ci, it.ci_lo, it.cin_total, it.ci0 = co_valid, ci_hi, ci_lo, cin_total, ci0
                            it.B, it.H, it.W = B, H, W
                            if chunks > 1:
                                it.tile_begin, it.tile_end = tiles * ch // chunks, tiles * (ch + 1) // chunks
                            # the bias gradient (sum of g over the pixels) rides with the conv's first block: the CTA has the
                            # gradient tile in shared memory anyway (a separate bias kernel took 3.7 of the step's 26 ms)
                            it.db = self.flat.data_ptr() + 4 * (b_off + co0) if (bi == 0 and not last) else 0
                            big.append(it)
                if not last:
                    continue
                sit = WgradSmallItem()                         # last conv: the bias sum of the fp32 gradient planes
                sit.g, sit.dw, sit.db = g.data_ptr(), dw, self.flat.data_ptr() + 4 * b_off
                sit.g32 = bp.g_y.data_ptr() if name == names[-1] else 0
                sit.s = 0
                sit.g_stride, sit.g_c0, sit.cout, sit.n_co = g.shape[-1], gc0, cout, n_co
                sit.s_channels, sit.s_c0, sit.n_c = 0, 0, 0
                sit.cin_total, sit.ci0, sit.B, sit.H, sit.W = cin_total, 0, B, H, W
                small.append(sit)
            big_arr = (WgradItem * max(1, len(big)))(*big)
            small_arr = (WgradSmallItem * max(1, len(small)))(*small)
            tcs.sort(key=lambda t_: -((t_.tile_end - t_.tile_begin) if t_.tile_end > 0 else t_.B * ((t_.H + 7) // 8) * ((t_.W + 15) // 16)))   # longest first
            tc_arr = (WgradTcItem * max(1, len(tcs)))(*tcs)
            tables.append((_struct_array_to_device(big_arr, WgradItem, self.dev) if big else None, len(big),
                           _struct_array_to_device(small_arr, WgradSmallItem, self.dev) if small else None, len(small),
                           max([s_.B * s_.H for s_ in small] + [1]),
                           _struct_array_to_device(tc_arr, WgradTcItem, self.dev) if tcs else None, len(tcs)))
        self._tc_maps = torch.frombuffer(bytearray(b"".join(maps_raw)), dtype=torch.uint8).to(self.dev) if maps_raw else None
        self._tables = {key: tables}
        return tables

    # ------------------------------------------------------------------ backward
    def backward(self, grad_fake_H, all_reduce=True):
        """dL/dW and dL/db of every generator conv for the gradient w.r.t. the last forward's output; with an initialised
        process group the result is averaged over the ranks.  Returns the gradient w.r.t. model_input."""
        if self._state is None:
            raise capi.EsrError("GeneratorTrainer.backward before forward")
        plan, filters, margin = self._state
        self._state = None
        l = capi.lib()
        world = dist.get_world_size(self.group) if (all_reduce and dist.is_available() and dist.is_initialized()) else 1
        with torch.cuda.device(self.dev):
            cur = torch.cuda.current_stream()
            bp = self._backward_plan(plan)
            gout = grad_fake_H.contiguous().float()
            if filters is None:
                bp.g_y.copy_(gout)                             # the last conv's bias gradient reads the fp32 planes from bp.g_y
            g_in = generator_backward_eager(plan, bp, filters, margin, gout)
            tables = self._items(plan, bp)
            for src16, dst in self._bf16_copies:
                dst.copy_(src16)
            self.flat.zero_()                                  # chunked high-resolution items accumulate
            handles = []
            def launch(tab):
                big, nbig, small, nsmall, max_rows, tcs, ntc = tab
                if ntc:
                    capi.check(l.esr_wgrad_tc(C.c_void_p(tcs.data_ptr()), ntc, C.c_void_p(self._tc_maps.data_ptr()), capi.stream_ptr()))
                if nbig:
                    capi.check((l.esr_wgrad16r if self.split_k else l.esr_wgrad16)(C.c_void_p(big.data_ptr()), nbig, capi.stream_ptr()))
                if nsmall:
                    capi.check(l.esr_wgrad_small(C.c_void_p(small.data_ptr()), nsmall, max_rows, capi.stream_ptr()))

            def exchange(lo, hi):
                ev = torch.cuda.Event()
                ev.record(cur)
                with torch.cuda.stream(self.comm):
                    self.comm.wait_event(ev)
                    return dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            handles = []
            for (a, b_), tab in zip(zip(self._group_cuts[:-1], self._group_cuts[1:]), tables):
                launch(tab)
                if world > 1:                                  # this group's exchanges run under the next group's kernels
                    handles += [exchange(lo, hi) for (_, lo, hi) in self.buckets[a:b_]]
            for hnd in handles:
                hnd.wait()
            if world > 1:
                cur.wait_stream(self.comm)
        return g_in

    def flat_parameter(self):
        """ONE fp32 parameter holding every generator weight in the layout of the flat gradient buffer (its ``.grad``):
        each ``p.data`` becomes a view into it, so module, state_dict and checkpoints are unchanged, while an optimiser
        built on ``[trainer.flat_parameter()]`` updates the whole network with one fused launch (torch's fused Adam over the
        702 separate tensors is 20 launches of ~40 CTAs: 0.78 ms of a 9 ms step)."""
        if getattr(self, "_flat_param", None) is None:
            w = torch.zeros_like(self.flat)
            for key, (off, n, p) in self.slices.items():
                w[off:off + n].copy_(p.data.reshape(-1))
                p.data = w[off:off + n].view_as(p)
            self._flat_param = torch.nn.Parameter(w)
            self._flat_param.grad = self.flat
            self.G.weights_changed()
        return self._flat_param

    def grads_like(self, params):
        """Copies of the flat gradient buffer's slices for `params` (for autograd, which may keep what it is handed)."""
        by_id = {id(p): (off, n) for (off, n, p) in self.slices.values()}
        flat = self.flat.clone()
        return [flat[by_id[id(p)][0]:by_id[id(p)][0] + by_id[id(p)][1]].view_as(p) if id(p) in by_id else None for p in params]

    def grad_bytes(self):
        return self.flat.numel() * 4


class GanTrainer:
    """One iteration of the reference's GAN training step (codes/models/SRRaGAN_model.py:307-547) in the shipped
    configuration (options/train/train_esrgan_CEM.json: ``gan_type`` wgan-gp, non-relativistic critic, l1 pixel
    criterion, range loss), data parallel, one process per GPU:

      fake_H = netG(model_input)                                   :349    this package's kernels, activations kept
      critic:  l_d = (2*(-D(real).mean()) + 2*D(fake.detach()).mean())/2   :378-388
               + gp_weight * mean((|d D(interp) / d interp|_2 - 1)^2), interp = u*fake + (1-u)*real   :390-399, loss.py:244-263
               l_d.backward(); optimizer_D.step()                  :433,444
      generator: l_g = pixel_weight*L1(fake_H, HR) + range_weight*RangeLoss(fake_H) + gan_weight*(-D(fake_H).mean())
               l_g.backward(); optimizer_G.step()                  :478-533,565     dgrad + weight-gradient kernels

    The critic (esr_b200.discriminator, 1.6 % of the step's FLOPs, needs a double backward) is torch code on torch autograd;
    its gradient is flattened into one buffer and averaged with ONE all_reduce (13.6 MB), the generator's 68.2 MB gradient
    goes out in buckets under the weight-gradient kernels (GeneratorTrainer.backward).  ``crop``: the CEM margin the
    reference removes from fake_H / HR before the losses (``HR_unpadder``, :343-355); 0 keeps the whole patch (BASELINE
    config 5's 128x128 patches: the critic's 8x8 head needs >= 64 pixels)."""

    def __init__(self, netG, netD, lr_G=1e-5, lr_D=1e-5, betas_G=(0.9, 0.999), betas_D=(0.9, 0.999), pixel_weight=0.0, gan_weight=1.0,
                 gp_weight=10.0, range_weight=5000.0, crop=0, process_group=None):
        self.gen = GeneratorTrainer(netG, process_group=process_group)
        self.netD = netD
        self.group = process_group
        self.w = dict(pix=float(pixel_weight), gan=float(gan_weight), gp=float(gp_weight), range=float(range_weight))
        self.crop = int(crop)
        d_params = [p for p in netD.parameters()]
        self.d_flat = torch.zeros(sum(p.numel() for p in d_params), dtype=torch.float32, device=self.gen.dev)
        off = 0
        for p in d_params:                                   # p.grad of every critic parameter is a view of one buffer
            p.grad = self.d_flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.d_params = d_params
        self.optimizer_G = torch.optim.Adam([self.gen.flat_parameter()], lr=lr_G, betas=betas_G, fused=True)   # one launch
        self.optimizer_D = torch.optim.Adam(d_params, lr=lr_D, betas=betas_D, fused=True)
        self.log = {}
        # the torch halves of the step (critic update; generator loss terms through the critic) replay as CUDA graphs
        self.use_graphs = os.environ.get("ESR_GAN_GRAPHS", "1") != "0"
        self.channels_last = os.environ.get("ESR_GAN_CHANNELS_LAST", "1") != "0"
        self._graphs = {}

    def _cropped(self, t):
        c = self.crop
        return t[..., c:t.size(-2) - c, c:t.size(-1) - c] if c else t

    # ---- the two torch halves of the step, written once and run either eagerly or as captured CUDA graphs
    def _critic_losses(self, real, fake_d, u):
        """Critic forward on real / fake / interpolates and the backward of its loss into d_flat (no optimiser step)."""
        w, D = self.w, self.netD
        self.d_flat.zero_()
        if self.channels_last:                               # the critic's activations in NHWC: cuDNN's tensor-core kernels take
            real = real.contiguous(memory_format=torch.channels_last)      # them as they are (its NCHW <-> NHWC conversion kernels
            fake_d = fake_d.contiguous(memory_format=torch.channels_last)  # were ~200 launches of the eager step)
        pred_real, pred_fake = D(real), D(fake_d)
        l_d_real, l_d_fake = -2.0 * pred_real.mean(), 2.0 * pred_fake.mean()
        interp = (u * fake_d + (1 - u) * real).requires_grad_(True)
        crit = D(interp)
        g_interp, = torch.autograd.grad(crit, interp, torch.ones_like(crit), create_graph=True)
        l_d_gp = w['gp'] * ((g_interp.reshape(g_interp.size(0), -1).norm(2, dim=1) - 1) ** 2).mean()
        ((l_d_real + l_d_fake) / 2 + l_d_gp).backward()
        return {'l_d_real': l_d_real.detach(), 'l_d_fake': l_d_fake.detach(), 'l_d_gp': l_d_gp.detach(),
                'D_real': pred_real.detach().mean(), 'D_fake': pred_fake.detach().mean()}

    def _generator_losses(self, fake_full, real):
        """Generator loss terms on fake_H and their gradient w.r.t. it (through the fixed critic)."""
        w, D = self.w, self.netD
        fake = self._cropped(fake_full)
        log, l_g = {}, 0
        if w['pix']:
            l_pix = (fake - real).abs().mean()
            l_g = l_g + w['pix'] * l_pix
            log['l_g_pix'] = l_pix.detach()
        if w['range']:                                       # mean excursion out of [0, 1] (loss.py:236-242)
            l_range = torch.maximum(fake - 1, -fake).clamp_min(0).mean()
            l_g = l_g + w['range'] * l_range
            log['l_g_range'] = l_range.detach()
        l_gan = -w['gan'] * D(fake.contiguous(memory_format=torch.channels_last) if self.channels_last else fake).mean()
        log['l_g_gan'] = l_gan.detach()
        grad, = torch.autograd.grad(l_g + l_gan, fake_full)
        return grad, log

    def _captured(self, key, fn, static_inputs):
        """CUDA graph of fn(*static_inputs) (torch autograd included: the critic is ~600 small kernels whose launches, not
        their run time, bound the eager step).  Warm-up runs on a side stream with the critic's BatchNorm statistics restored
        afterwards; returns (graph, static inputs, static outputs)."""
        ent = self._graphs.get(key)
        if ent is None:
            bufs = [(b, b.detach().clone()) for b in self.netD.buffers()]
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(3):
                    fn(*static_inputs)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            for b, saved in bufs:
                b.copy_(saved)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = fn(*static_inputs)
            ent = self._graphs[key] = (g, static_inputs, out)
        return ent

    def step(self, model_input, var_H, generator_step=True, interpolation=None):
        world = dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1
        real = self._cropped(var_H).contiguous()
        fake_full = self.gen.forward(model_input)            # leaf; the graph below it is this package's backward
        fake_d = self._cropped(fake_full).detach()
        u = torch.rand(real.size(0), 1, 1, 1, device=real.device) if interpolation is None else interpolation
        # ---- critic
        for p in self.d_params:
            p.requires_grad_(True)
        if self.use_graphs:
            key = ('critic', tuple(real.shape))
            if key not in self._graphs:
                self._captured(key, self._critic_losses, (real.clone(), fake_d.clone(), u.clone()))
            g, (s_real, s_fake, s_u), out = self._graphs[key]
            s_real.copy_(real), s_fake.copy_(fake_d), s_u.copy_(u)
            g.replay()
            self.log = {k: v.clone() for k, v in out.items()}
        else:
            self.log = self._critic_losses(real, fake_d, u)
        if world > 1:
            dist.all_reduce(self.d_flat, op=dist.ReduceOp.AVG, group=self.group)
        self.optimizer_D.step()
        if not generator_step:
            self.gen._state = None
            return self.log
        # ---- generator (the critic is a fixed function here: its parameters take no gradient, :465-467)
        for p in self.d_params:
            p.requires_grad_(False)
        if self.use_graphs:
            key = ('generator', tuple(fake_full.shape))
            if key not in self._graphs:
                self._captured(key, self._generator_losses, (fake_full.detach().clone().requires_grad_(True), real.clone()))
            g, (s_fake, s_real), (grad, out) = self._graphs[key]
            with torch.no_grad():
                s_fake.copy_(fake_full), s_real.copy_(real)
            g.replay()
            self.log.update({k: v.clone() for k, v in out.items()})
        else:
            grad, out = self._generator_losses(fake_full, real)
            self.log.update(out)
        self.gen.backward(grad)                              # dgrad chain, weight gradients, bucketed all-reduce
        self.optimizer_G.step()
        return self.log

    def grad_bytes(self):
        return self.gen.grad_bytes(), self.d_flat.numel() * 4
