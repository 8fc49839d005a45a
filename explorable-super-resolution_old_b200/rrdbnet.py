"""RRDBNet generator, host side.  Drop-in for ``models.modules.architecture.RRDBNet``
(codes/models/modules/architecture.py:102-175): same constructor, attributes and state_dict
keys; ``forward`` runs the recorded sm_100a kernel sequence instead of 351 aten convs.

The module tree below only holds parameters under the reference's names
(``model.0``, ``model.1.sub.{r}.RDB{d}.convs.{i}.0``, ``model.1.sub.{nb}``, ``model.{2,3}.1``,
``model.4``, ``model.6``; latent input channels first on dim 1) so that checkpoints and
``process_loaded_state_dict`` (codes/models/base_model.py:113-144) keep working.
"""
import math
import threading
import weakref

import torch
import torch.nn as nn

from . import _capi as capi
from .engine import GEngine, GPlan


_CACHE_LOCK = threading.RLock()      # engine / plan caches of every RRDBNet are built under it (module level: an nn.Module
                                     # must stay deep-copyable, a lock attribute is not); plans are keyed per host thread


def _conv(cin, cout):
    return nn.Conv2d(cin, cout, 3, 1, 1, bias=True)


class _DenseBlock(nn.Module):            # parameter holder for ResidualDenseBlock_5C (block.py:196-242)
    def __init__(self, nf, gc, nz):
        super().__init__()
        self.convs = nn.ModuleList(
            [nn.Sequential(_conv(nf + i * gc + nz, gc), nn.LeakyReLU(0.2, True)) for i in range(4)] +
            [nn.Sequential(_conv(nf + 4 * gc + nz, nf))])


class _RRDB(nn.Module):                  # parameter holder for RRDB (block.py:245-270)
    def __init__(self, nf, gc, nz):
        super().__init__()
        self.num_latent_channels = nz
        self.RDB1, self.RDB2, self.RDB3 = _DenseBlock(nf, gc, nz), _DenseBlock(nf, gc, nz), _DenseBlock(nf, gc, nz)


class _Trunk(nn.Module):                 # parameter holder for ShortcutBlock (block.py:76-97)
    def __init__(self, nf, gc, nb, nz):
        super().__init__()
        self.num_latent_channels = nz
        self.sub = nn.ModuleList([_RRDB(nf, gc, nz) for _ in range(nb)] + [_conv(nf + nz, nf)])


class RRDBNet(nn.Module):
    def __init__(self, in_nc, out_nc, nf, nb, gc=32, upscale=4, norm_type=None, act_type='leakyrelu', mode='CNA',
                 upsample_mode='upconv', latent_input=None, num_latent_channels=None):
        super().__init__()
        if nf != 64:
            raise NotImplementedError("nf=%d: the kernels are specialised for nf=64, gc=32" % nf)
        if norm_type is not None or act_type != 'leakyrelu' or mode != 'CNA' or upsample_mode != 'upconv':
            raise NotImplementedError("only norm_type=None, leakyrelu, mode='CNA', upconv are built (the production config)")
        if latent_input is not None and 'HR_rearranged' in latent_input and 'all_layers' in latent_input:
            # the reference's own forward raises 'Unsupported yet' for all_layers_HR_rearranged (architecture.py:167-169)
            raise NotImplementedError("latent_input all_layers_HR_rearranged is unsupported (in the reference too)")
        if latent_input is not None and not any(d in latent_input for d in ('HR_downscaled', 'HR_rearranged', 'LR')):
            raise NotImplementedError("latent_input %r: expected <all_layers|first_layer>_<HR_downscaled|HR_rearranged|LR>" % (latent_input,))
        self.latent_input = latent_input
        # 'LR' domain (architecture.py:159,165-166): Z has the LR image's size and is assigned to ``self.Z`` by the caller
        # (SRRaGAN_model.py:260-261); forward(x) takes the 3 image channels only.  See _lr_domain_input.
        self.latent_domain = None if latent_input is None else next(d for d in ('HR_downscaled', 'HR_rearranged', 'LR') if d in latent_input)
        # 'HR_rearranged' (first_layer only): Z [B, Cz*sf^2, H, W] = the HR latent rearranged into channels, assigned to
        # ``.Z`` like an LR-domain latent; it enters the first conv only (architecture.py:109-110: num_latent_channels *= sf^2)
        z_rearranged = 0
        if self.latent_domain == 'HR_rearranged' and num_latent_channels:
            num_latent_channels = num_latent_channels * upscale ** 2
            z_rearranged = num_latent_channels
        nz_in = num_latent_channels if (latent_input is not None and num_latent_channels) else 0
        if z_rearranged:
            nz_in = 0                                  # the engine's head kernels un-view an HR-domain latent: not this one
        self.num_latent_channels = 1 * num_latent_channels if num_latent_channels is not None else None
        self.upscale = upscale
        all_layers = latent_input is not None and 'all_layers' in latent_input
        nz = nz_in if all_layers else 0
        gc = 32                                        # hard-coded in the reference (architecture.py:123)
        n_up = 1 if upscale == 3 else int(math.log(upscale, 2))
        mods = [_conv(in_nc + nz_in + z_rearranged, nf), _Trunk(nf, gc, nb, nz)]
        for _ in range(n_up):
            mods.append(nn.Sequential(nn.Upsample(scale_factor=3 if upscale == 3 else 2, mode='nearest'),
                                      _conv(nf, nf), nn.LeakyReLU(0.2, True)))
        mods += [_conv(nf + nz, nf), nn.LeakyReLU(0.2, True), _conv(nf + nz, out_nc)]
        self.model = nn.ModuleList(mods)
        if self.latent_domain == 'LR' and nz:
            self.latent_upsampler = nn.Upsample(scale_factor=upscale if upscale == 3 else 2)   # attribute parity (:137-139)
        self._cfg = dict(nb=nb, nz_in=nz_in, all_layers=all_layers, out_nc=out_nc, in_nc=in_nc, upscale=upscale, z_rearranged=z_rearranged)
        self._engine, self._engine_key, self._plans = None, None, {}
        self._packed_params, self._dgrad, self._bplans = None, None, {}
        self.precise_outer = True
        self.outer_mode = None                         # None: engine default ("f16"); "split" | "bf16" for experiments
        self.debug_simt = False
        # Replay the differentiable forward and its data-gradient backward as CUDA graphs (one host launch each instead
        # of ~360).  Z_optimizer switches it on for its loop; shapes, margin and CEM filters key the captured graphs.
        self.use_cuda_graphs = False

    # ------------------------------------------------------------------ engine plumbing
    def _named_convs(self):
        return {n[:-len(".weight")]: None for n, _ in self.named_parameters() if n.endswith(".weight")}

    def engine(self):
        with _CACHE_LOCK:
            return self._engine_locked()

    def _engine_locked(self):
        params = dict(self.named_parameters())
        ptr_key = tuple(p.data_ptr() for p in params.values()) + (self.precise_outer, self.outer_mode)
        key = ptr_key + tuple(p._version for p in params.values())
        if self._engine is not None and key == self._engine_key:
            return self._engine
        same_storage = self._engine is not None and getattr(self, "_engine_ptr_key", None) == ptr_key
        eng = self._engine if same_storage else GEngine(precise_outer=self.precise_outer, outer_mode=self.outer_mode, **self._cfg)
        packed = {}
        for name in eng.convs:
            w, b = params[name + ".weight"], params[name + ".bias"]
            if not w.is_cuda:
                raise capi.EsrError("RRDBNet parameters live on %s; move the module to a B200 (no CPU path)" % w.device)
            packed[name] = (w.detach().contiguous().float(), b.detach().contiguous().float())
        eng.pack(packed)
        if same_storage:
            # a weight update in place (optimizer step, load_state_dict): the packed images are refreshed inside their
            # existing buffers, so plans, recorded launch sequences and captured graphs stay valid
            self._packed_params = packed
            if self._dgrad is not None:
                self._dgrad.pack(packed)
        else:
            self._engine, self._plans = eng, {}
            self._packed_params, self._dgrad, self._bplans = packed, None, {}
        self._engine_key, self._engine_ptr_key = key, ptr_key
        return self._engine

    def weights_changed(self):
        """Forces the next forward to re-pack the weights.  In-place updates are normally seen through the parameters'
        version counters, but torch's fused optimizers (``Adam(fused=True)``: ``_fused_adam_``) write the parameters
        without bumping them, so the training paths (training.GeneratorTrainer.forward, _TrainFn) call this every step."""
        with _CACHE_LOCK:
            self._engine_key = None

    def backward_plan(self, plan):
        with _CACHE_LOCK:
            return self._backward_plan_locked(plan)

    def _backward_plan_locked(self, plan):
        from .backward import DgradSpecs, BackwardPlan
        if self._dgrad is None:
            self._dgrad = DgradSpecs(self._engine)
            self._dgrad.pack(self._packed_params)
        key = id(plan)
        if key not in self._bplans:
            if len(self._bplans) >= 2:
                self._bplans.pop(next(iter(self._bplans)))
            self._bplans[key] = (plan, BackwardPlan(plan, self._dgrad, use_simt=self.debug_simt))
        return self._bplans[key][1]

    def plan(self, B, h, w, m, keep, slot=0):
        """Buffers + recorded launches for one geometry.  `slot` selects an independent buffer set, for callers
        that keep several sub-batches in flight on different streams."""
        # A plan's buffers hold one forward's activations: two host threads driving the same module (the
        # module must be replica-safe) get their own plans; an evicted plan stays alive as long as a captured graph or an
        # autograd context still references it (they hold the object, the cache only holds the most recent four per thread).
        tid = threading.get_ident()
        thread_key = 0 if tid == threading.main_thread().ident else tid
        with _CACHE_LOCK:
            eng = self._engine_locked()
            key = (B, h, w, m, keep, self.debug_simt, slot, thread_key)
            if key not in self._plans:
                mine = [k for k in self._plans if k[-1] == thread_key]
                if len(mine) >= 4:
                    self._plans.pop(mine[0])
                dev = next(self.parameters()).device
                self._plans[key] = GPlan(eng, B, h, w, m, dev, keep_activations=keep, use_simt=self.debug_simt)
            return self._plans[key]

    def forward(self, x):
        return run_generator(self, x, margin=0, cem_filters=None)

    def train(self, mode=True):
        # switching between training and evaluation re-packs the weights on the next forward (see weights_changed)
        self._engine_key = None
        return super().train(mode)


def _capture(run, device):
    """Runs `run` once on a side stream (lazy initialisation inside the library happens outside the capture), then
    captures it.  Returns (graph, value returned by the captured call)."""
    with torch.cuda.device(device):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.current_stream().synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            res = run()
    return graph, res


class _GraphedStep:
    """Forward (input plumbing, 351 convs, CEM projection) and backward (CEM adjoint, 351 dgrads, input adjoint) of one
    plan as two CUDA graphs over fixed buffers.  Both are captured on the first differentiable forward, on the calling
    thread (autograd runs backward hooks on a worker thread, where a capture would be fragile)."""

    def __init__(self, net, plan, margin, cem_filters, x):
        from .backward import generator_backward_eager
        B, sf, dev = plan.B, net.upscale, x.device
        crop = sf * margin
        H4, W4 = sf * plan.hp, sf * plan.wp
        onc = plan.y.size(1)
        self.plan, self.filters = plan, cem_filters
        self.x = torch.empty_like(x)
        self.x.copy_(x)
        self.out = torch.empty(B, onc, H4 - 2 * crop, W4 - 2 * crop, device=dev, dtype=torch.float32)
        self.ws = torch.empty(max(1, 2 * B * onc * plan.hp * plan.wp), device=dev, dtype=torch.float32)
        self.g = torch.zeros_like(self.out)
        self.bp = net.backward_plan(plan) if net._cfg["nz_in"] > 0 else None

        def fwd():
            _forward_eager(plan, self.x, cem_filters, crop, self.out, self.ws)
        self.fwd_graph, _ = _capture(fwd, dev)
        self.bwd_graph, self.g_in = (None, None)
        if self.bp is not None:
            self.bwd_graph, self.g_in = _capture(lambda: generator_backward_eager(plan, self.bp, cem_filters, margin, self.g), dev)

    def forward(self, x):
        self.x.copy_(x)
        self.fwd_graph.replay()
        return self.out.clone()

    def backward(self, g):
        self.g.copy_(g)
        self.bwd_graph.replay()
        return self.g_in.clone()


def _forward_eager(plan, x, cem_filters, crop, out, ws):
    y = plan.run_g(x)
    if cem_filters is None:
        out.copy_(y)
    else:
        B, onc = y.size(0), y.size(1)
        capi.cem_call("project", cem_filters, capi.ptr(y), capi.ptr(plan.lr_pad), B, onc, y.size(2), y.size(3), crop,
                      capi.ptr(out), capi.ptr(ws), capi.stream_ptr())


class _GeneratorFn(torch.autograd.Function):
    """G (+ optional CEM projection) as one autograd node: data gradient only.

    The activations the backward needs (the sign of the 279 LeakyReLU outputs) stay in the plan's buffers, which are
    shared by every forward of the same geometry: each differentiable forward stamps the plan, and a backward whose
    stamp is no longer the plan's latest raises instead of silently differentiating through another input's
    activations (two forwards before one backward, or a second backward after another forward).  The gradient is
    defined for the latent channels only: the LR channels of d(model_input) are zero (the reference's Z optimisation
    never asks for them, Z_optimization.py:545-553)."""

    @staticmethod
    def forward(ctx, x, net, margin, cem_filters, need_grad):
        B, C, h, w = x.shape
        plan = net.plan(B, h, w, margin, keep=need_grad)
        sf = net.upscale
        ctx.plan, ctx.cem_filters, ctx.margin, ctx.net, ctx.graphed = plan, cem_filters, margin, net, None
        if need_grad:
            plan.fwd_stamp = getattr(plan, "fwd_stamp", 0) + 1
        ctx.stamp = getattr(plan, "fwd_stamp", 0)
        with torch.cuda.device(x.device):
            if need_grad and net.use_cuda_graphs:
                key = (id(cem_filters), margin)
                if key not in plan.graphed:
                    plan.graphed[key] = _GraphedStep(net, plan, margin, cem_filters, x)
                ctx.graphed = plan.graphed[key]
                return ctx.graphed.forward(x)
            crop = sf * margin
            onc = plan.y.size(1)
            out = torch.empty(B, onc, sf * plan.hp - 2 * crop, sf * plan.wp - 2 * crop, device=x.device, dtype=torch.float32)
            ws = torch.empty(2 * B * onc * plan.hp * plan.wp, device=x.device, dtype=torch.float32) if cem_filters is not None else None
            _forward_eager(plan, x, cem_filters, crop, out, ws)
        return out

    @staticmethod
    def backward(ctx, g):
        from .backward import generator_backward
        if getattr(ctx.plan, "fwd_stamp", 0) != ctx.stamp:
            raise capi.EsrError("backward through a generator forward whose activations were overwritten by a later "
                                "differentiable forward of the same shape (call backward before the next forward)")
        if ctx.graphed is not None and ctx.graphed.bwd_graph is not None:
            with torch.cuda.device(g.device):
                return ctx.graphed.backward(g.contiguous().float()), None, None, None, None
        return generator_backward(ctx, g), None, None, None, None


def capture_inference(net, x_static, margin, cem_filters, slot=0):
    """CUDA graph of the whole inference forward (input plumbing, the 351 convs, CEM projection) reading the fixed
    device buffer `x_static` and writing a fixed output buffer.  Returns (graph, out).  Replaying it costs one launch
    on the host; `slot` selects a private set of plan buffers so that several graphs can be in flight."""
    if not x_static.is_cuda or x_static.dtype != torch.float32 or not x_static.is_contiguous():
        raise capi.EsrError("capture_inference: expected a contiguous float32 CUDA tensor")
    capi.require_device(x_static.device.index if x_static.device.index is not None else torch.cuda.current_device())
    B, C, h, w = x_static.shape
    sf = net.upscale
    if C != net._cfg["nz_in"] * sf * sf + 3:
        raise ValueError("expected %d input channels, got %d" % (net._cfg["nz_in"] * sf * sf + 3, C))
    plan = net.plan(B, h, w, margin, keep=False, slot=slot)
    crop = sf * margin
    H4, W4 = sf * plan.hp, sf * plan.wp
    onc = plan.y.size(1)
    out = torch.empty(B, onc, H4 - 2 * crop, W4 - 2 * crop, device=x_static.device, dtype=torch.float32)
    ws = torch.empty(2 * B * onc * plan.hp * plan.wp, device=x_static.device, dtype=torch.float32) if cem_filters is not None else None

    def run():
        y = plan.run_g(x_static)
        if cem_filters is None:
            out.copy_(y)
        else:
            capi.cem_call("project", cem_filters, capi.ptr(y), capi.ptr(plan.lr_pad), B, onc, H4, W4, crop, capi.ptr(out),
                          capi.ptr(ws), capi.stream_ptr())
    with torch.cuda.device(x_static.device):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run()                                      # warm-up outside the capture (lazy initialisation in the library)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.current_stream().synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            run()
    graph._esr_keep = (plan, ws, x_static)
    return graph, out


_TRAINERS = weakref.WeakKeyDictionary()      # RRDBNet -> its GeneratorTrainer (kept off the module: deepcopy / state_dict stay plain)


class _TrainFn(torch.autograd.Function):
    """G (+ CEM) with trainable parameters as one autograd node.  forward keeps the activations
    (training.GeneratorTrainer.forward); backward runs the dgrad chain and the weight-gradient kernels of csrc/wgrad.cu
    and returns dL/dW, dL/db of the 351 convs to autograd, which accumulates them into ``p.grad`` as for any torch
    module (gradient accumulation over several backward calls, ``optimizer.zero_grad()``, DDP hooks all behave as
    usual).  With an initialised process group and ``net.esr_all_reduce`` (default True) the gradients are averaged over
    the ranks inside backward, bucketed under the weight-gradient kernels (training.py).  One forward per backward:
    the activations live in the plan's buffers (same rule as _GeneratorFn)."""

    @staticmethod
    def forward(ctx, x, net, margin, cem_filters, *params):
        from .training import GeneratorTrainer
        with _CACHE_LOCK:
            tr = _TRAINERS.get(net)
            if tr is None or tr.dev != x.device:
                tr = _TRAINERS[net] = GeneratorTrainer(net, attach_grads=False)
        out = tr.forward(x, margin=margin, filters=cem_filters, leaf=False)
        ctx.trainer, ctx.stamp, ctx.net = tr, tr.stamp, net
        return out

    @staticmethod
    def backward(ctx, g):
        tr = ctx.trainer
        if tr._state is None or tr.stamp != ctx.stamp:
            raise capi.EsrError("backward through a training forward of the generator whose activations were overwritten "
                                "by a later forward (or a second backward): one forward per backward")
        params = list(ctx.net.parameters())
        g_in = tr.backward(g, all_reduce=getattr(ctx.net, "esr_all_reduce", True))
        grads = tr.grads_like(params)
        grads = [gr if need else None for gr, need in zip(grads, ctx.needs_input_grad[4:])]
        return (g_in if ctx.needs_input_grad[0] else None, None, None, None) + tuple(grads)


def _lr_domain_input(net, x, margin):
    """LR-domain latent (architecture.py:159-166): ``x`` is the bare LR image, ``net.Z`` [B, Cz, H, W] has the size of the
    image the generator sees (in eval mode the CEM wrapper replicate-pads x by its margin first, CEMnet.py:180-181, so Z
    must already have the padded size — as in the reference, where a mismatch fails in torch.cat).  The reference
    concatenates Z to the LR-resolution convs as it is and its nearest-neighbour upsampling (``latent_upsampler`` once
    per upconv stage) to the HR convs.  That is exactly the HR_downscaled path fed with Z_HR = nearest_upsample(Z, sf):
    the centre-2x2 mean that path takes of every sf x sf block returns Z bit-exactly.  Plumbing in torch (pad, repeat,
    view, cat; differentiable w.r.t. Z and indexing-only), arithmetic in the kernels."""
    Z = getattr(net, 'Z', None)
    nz_in, sf = net._cfg["nz_in"], net.upscale
    if Z is None:
        raise AttributeError("RRDBNet with an LR-domain latent input: assign the latent to .Z before forward (SRRaGAN_model.py:260-261)")
    if margin > 0:
        x = torch.nn.functional.pad(x.float(), (margin,) * 4, mode='replicate')
    if x.size(1) != 3 or Z.dim() != 4 or Z.size(1) != nz_in or Z.shape[0] != x.shape[0] or Z.shape[2:] != x.shape[2:]:
        raise RuntimeError("LR-domain latent: Z %s does not match the (padded) LR input %s with %d latent channels"
                           % (tuple(Z.shape), tuple(x.shape), nz_in))
    B, _, H, W = x.shape
    z_hr = Z.to(x.device).float().repeat_interleave(sf, 2).repeat_interleave(sf, 3)
    return torch.cat([z_hr.contiguous().view(B, nz_in * sf * sf, H, W), x.float()], 1)


def run_generator(net, x, margin, cem_filters):
    if not x.is_cuda:
        raise capi.EsrError("RRDBNet.forward: expected a CUDA tensor; this package has no CPU or PyTorch fallback")
    if net.latent_domain == 'HR_rearranged':
        return _run_rearranged(net, x, margin, cem_filters)
    if net.latent_domain == 'LR':
        crop = net.upscale * margin
        out = _run_generator(net, _lr_domain_input(net, x, margin), 0, cem_filters)
        return out[..., crop:out.size(-2) - crop, crop:out.size(-1) - crop] if crop else out
    return _run_generator(net, x, margin, cem_filters)


def _run_rearranged(net, x, margin, cem_filters):
    """first_layer_HR_rearranged (architecture.py:109-110,159): ``net.Z`` [B, Cz*sf^2, H, W] is concatenated in front of the
    (padded) image for the first conv only.  Forward only: no gradient w.r.t. Z or the weights is built for this variant."""
    Z = getattr(net, 'Z', None)
    zr, sf = net._cfg["z_rearranged"], net.upscale
    if Z is None:
        raise AttributeError("RRDBNet with an HR_rearranged latent input: assign the latent to .Z before forward (SRRaGAN_model.py:260-261)")
    if torch.is_grad_enabled() and (Z.requires_grad or x.requires_grad or any(p.requires_grad for p in net.parameters())):
        raise NotImplementedError("gradients through the first_layer_HR_rearranged variant are not built (forward only)")
    dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
    capi.require_device(dev)
    xp = torch.nn.functional.pad(x.float(), (margin,) * 4, mode='replicate') if margin > 0 else x.float()
    if xp.size(1) != 3 or Z.dim() != 4 or Z.size(1) != zr or Z.shape[0] != xp.shape[0] or Z.shape[2:] != xp.shape[2:]:
        raise RuntimeError("HR_rearranged latent: Z %s does not match the (padded) LR input %s with %d rearranged channels"
                           % (tuple(Z.shape), tuple(xp.shape), zr))
    B, _, H, W = xp.shape
    plan = net.plan(B, H, W, 0, keep=False)
    crop = sf * margin
    onc = plan.y.size(1)
    with torch.cuda.device(x.device):
        plan.run_prep_rearranged(xp.contiguous(), Z.to(x.device).float())
        plan.run_convs()
        y = plan.y
        out = torch.empty(B, onc, sf * H - 2 * crop, sf * W - 2 * crop, device=x.device, dtype=torch.float32)
        if cem_filters is None:
            out.copy_(y[..., crop:sf * H - crop, crop:sf * W - crop])
        else:
            ws = torch.empty(2 * B * onc * H * W, device=x.device, dtype=torch.float32)
            capi.cem_call("project", cem_filters, capi.ptr(y), capi.ptr(plan.lr_pad), B, onc, y.size(2), y.size(3), crop,
                          capi.ptr(out), capi.ptr(ws), capi.stream_ptr())
    return out


def _run_generator(net, x, margin, cem_filters):
    dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
    capi.require_device(dev)
    nz_in = net._cfg["nz_in"]
    sf = net.upscale
    if x.size(1) != nz_in * sf * sf + 3:
        raise ValueError("expected %d input channels (Z.view(B,%d,h,w) ++ LR), got %d" % (nz_in * sf * sf + 3, nz_in * sf * sf, x.size(1)))
    need_grad = torch.is_grad_enabled() and x.requires_grad
    trainable = any(p.requires_grad for p in net.parameters())
    if trainable and not torch.is_grad_enabled():
        # a validation forward between training steps: a fused optimiser may have written the weights since the last
        # training forward without bumping their versions (weights_changed), so re-pack rather than trust the cache key
        net.weights_changed()
    if torch.is_grad_enabled() and trainable:
        # the reference's training step (SRRaGAN_model.py:349,533: fake_H = netG(model_input) ... l_g_total.backward()):
        # data-gradient pass + weight-gradient kernels behind one autograd node
        params = [p for p in net.parameters()]
        return _TrainFn.apply(x.contiguous().float(), net, margin, cem_filters, *params)
    xc = x.contiguous().float()
    return _GeneratorFn.apply(xc, net, margin, cem_filters, need_grad)
