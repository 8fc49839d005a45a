"""Consistency Enforcing Module, host side.  Drop-in for the reference's ``CEM/CEMnet.py``:
``Get_CEM_Config``, ``CEMnet``, ``CEM_PyTorch``, ``Filter_Layer``, ``Return_kernel``,
``Adjust_State_Dict_Keys`` keep their names, arguments, attributes and state_dict keys
(codes/CEM/CEMnet.py:14-245), while every tensor operation runs in libesr_b200.so.

The fixed filters are derived once at construction time in float64 numpy (init-time host code,
codes/CEM/CEMnet.py:105-126 and codes/CEM/imresize_CEM.py:18-94); for the default bicubic
kernel (and a mildly blurred one) they are rank-1 and the CUDA kernels consume their 1-D factors;
estimated / user-supplied kernels and strong blurs run the general 2-D stencils (csrc/cem2d.cu).
"""
import collections

import numpy as np
import torch
import torch.nn as nn

from . import _capi as capi

NFFT_ADD = 36


# --------------------------------------------------------------------------- filters
def _cubic_weights_1d(sf):
    """1-D factor of the x`sf` bicubic interpolation kernel (Keys a=-0.75, what cv2.INTER_CUBIC
    applies to a delta image in imresize_CEM.py:88-94): samples at the 2*sf (even sf) or
    4*sf-1 (odd sf) output positions within distance < 2 of the input sample."""
    a = -0.75
    # output pixel i sits at (i + 0.5)/sf - 0.5 input pixels; enumerate offsets around the origin
    offs = (np.arange(-4 * sf, 4 * sf + 1) + 0.5) / sf - 0.5
    d = np.abs(offs)
    w = np.where(d <= 1, ((a + 2) * d - (a + 3)) * d * d + 1,
                 np.where(d < 2, ((a * d - 5 * a) * d + 8 * a) * d - 4 * a, 0.0))
    nz = np.nonzero(w)[0]
    return w[nz[0]:nz[-1] + 1]


def sampling_phase(sf):
    """(pre, post): zeros before / after each sample when zero-stuffing (imresize_CEM.py:73-86)."""
    post = sf // 2
    return sf - post - 1, post


def default_ds_kernel_1d(sf):
    """1-D factor d of the downscaling kernel: ds_kernel == outer(d, d) (CEMnet.py:218-219)."""
    pre, post = sampling_phase(sf)
    up = _cubic_weights_1d(sf)
    up = np.concatenate([np.zeros(max(0, post - pre)), up, np.zeros(max(0, pre - post))])
    return (up[::-1] / sf).astype(np.float32)


def _conv2_full(a, b):
    out = np.zeros((a.shape[0] + b.shape[0] - 1, a.shape[1] + b.shape[1] - 1))
    for i in range(b.shape[0]):
        for j in range(b.shape[1]):
            out[i:i + a.shape[0], j:j + a.shape[1]] += b[i, j] * a
    return out


def _gaussian_2d(sigma):
    """Normalised isotropic Gaussian holding 99% of its 1-D mass (imresize_CEM.py:101-109):
    half width = ceil(-ppf(0.005) * sigma), ppf(0.005) of the unit normal = -2.5758293035489004."""
    half = int(np.ceil(2.5758293035489004 * sigma))
    g = np.exp(-0.5 * (np.arange(-half, half + 1) / sigma) ** 2)
    g2 = np.outer(g, g)
    return g2 / g2.sum()


def _rint(v):
    return int(np.round(v))


def _center_mass(k, sf):
    """Re-centre a user-supplied (already rot180'd) kernel on its centre of mass by zero padding,
    then trim it to the frame holding 99% of its root-energy, keeping a side the x`sf` polyphase
    split accepts (imresize_CEM.py:114-160)."""
    if k.ndim != 2 or k.shape[0] != k.shape[1]:
        raise ValueError("Currently supporting only square kernels")
    n = k.shape[0]
    cols, rows = np.meshgrid(np.arange(n), np.arange(n))
    # first moments in flipped coordinates, 1-based (what conv2(grid, k, 'valid') + 1 evaluates)
    gx = float(np.sum(cols[::-1, ::-1] * k)) + 1
    gy = float(np.sum(rows[::-1, ::-1] * k)) + 1
    pad = {"x": 2 * (n / 2 - gx), "y": 2 * (n / 2 - gy)}
    before = {a: max(0.0, -pad[a]) for a in pad}
    after = {a: max(0.0, pad[a]) for a in pad}
    extra = float(np.round(abs(pad["y"])) - np.round(abs(pad["x"])))   # > 0: x needs more zeros to stay square
    if extra != 0:
        grow = "x" if extra > 0 else "y"
        # the rounding of the two pads decides which side takes the odd zero
        lean_after = (np.round(after[grow]) - after[grow]) - (np.round(before[grow]) - before[grow]) > 0
        lo, hi = int(np.floor(abs(extra) / 2)), int(np.ceil(abs(extra) / 2))
        before[grow] = _rint(before[grow]) + (lo if lean_after else hi)
        after[grow] = _rint(after[grow]) + (hi if lean_after else lo)
    k = np.pad(k, ((_rint(before["y"]), _rint(after["y"])), (_rint(before["x"]), _rint(after["x"]))), mode="constant")
    if k.shape[0] != k.shape[1]:
        raise ValueError("re-centring the kernel left it non-square")
    n = k.shape[0]
    total = np.sqrt(np.sum(k ** 2))
    frames = [1.0] + [np.sqrt(np.sum(k[f:-f, f:-f] ** 2)) / total for f in range(1, int(np.ceil(n / 2)))]
    below = [f for f, e in enumerate(frames) if e < 0.99]
    if not below:
        raise ValueError("kernel too concentrated to trim (no frame holds < 99% of its energy)")
    cut = [below[0], below[0]]
    side = 0
    while (n - sum(cut) - 1 + (sf + 1) % 2) % sf != 0:
        cut[side] -= 1
        side ^= 1
    if min(cut) <= 0:
        raise ValueError("kernel support too small to trim to a x%d-compatible size" % sf)
    k = k[cut[0]:n - cut[1], cut[0]:n - cut[1]]
    return k / np.sum(k)


def _upscale_kernel(sf, upscale_kernel):
    """imresize(None, [sf, sf], return_upscale_kernel=True, kernel=...) (imresize_CEM.py:18-47)."""
    pre, post = sampling_phase(sf)
    pad_front, pad_back = max(0, post - pre), max(0, pre - post)   # compensates the uneven phase of even factors
    if isinstance(upscale_kernel, np.ndarray):
        if abs(1 - np.sum(upscale_kernel)) >= np.finfo(np.float32).eps:
            raise ValueError("Supplied non-default kernel does not sum to 1")
        k = _center_mass(np.rot90(upscale_kernel.astype(np.float64), 2), sf) * sf ** 2
        if (k.shape[0] + pad_front + pad_back - 1) % sf != 0:
            raise ValueError("Convolution-invalidated size should be an integer multiplication of the scale factor")
    else:
        w = _cubic_weights_1d(sf)
        k = np.outer(w, w)
        if upscale_kernel is not None and "blurry_cubic" in upscale_kernel:
            k = _conv2_full(k, _gaussian_2d(float(upscale_kernel[len("blurry_cubic_"):])))
    return np.pad(k, ((pad_front, pad_back), (pad_front, pad_back)), mode="constant")


def Return_kernel(ds_factor, upscale_kernel=None):
    """ds_kernel (CEMnet.py:218-219).  upscale_kernel: None / 'cubic' / 'reset_2_default' (bicubic),
    'blurry_cubic_<sigma>', or a square ndarray downscaling kernel summing to 1."""
    sf = int(ds_factor)
    if isinstance(upscale_kernel, str) and not any(w in upscale_kernel for w in ("cubic", "reset_2_default")):
        raise ValueError("unknown CEM kernel %r" % (upscale_kernel,))
    return (np.rot90(_upscale_kernel(sf, upscale_kernel), 2) / sf ** 2).astype(np.float32)


def _response_margin(resp, limit):
    """Depth of the border band where a filter's response to a constant image deviates by more than
    `limit` (CEMnet.py:28-42)."""
    n = resp.shape[0]
    r = resp / resp[n // 2, n // 2]
    r = np.where(r <= 0, limit / 2, r)
    bad = np.exp(-np.abs(np.log(r))) < limit
    col = np.flatnonzero(bad[:n // 2, n // 2])
    row = np.flatnonzero(bad[n // 2, :n // 2])
    return int(max(col[-1] + 1, row[-1] + 1))


def _conv2_same_ones(n, k):
    """conv2(ones(n,n), k, 'same') via summed tap coverage (k square, odd)."""
    full = _conv2_full(np.ones((n, n)), k)
    o = k.shape[0] // 2
    return full[o:o + n, o:o + n]


def Get_CEM_Config(sf):
    class config:
        scale_factor = sf
        desired_inv_hTh_energy_portion = 1 - 1e-6
        filter_pertubation_limit = 0.999
        lower_magnitude_bound = 0.01
    return config


def _rank1_factor(k2d):
    """f with k2d == outer(f, f) to 1e-6 of its peak, or None when the filter is not symmetric rank-1."""
    u, s, vt = np.linalg.svd(k2d.astype(np.float64))
    f = u[:, 0] * np.sqrt(s[0])
    g = vt[0] * np.sqrt(s[0])
    if f[np.argmax(np.abs(f))] < 0:
        f, g = -f, -g
    if np.abs(np.outer(f, g) - k2d).max() > 1e-6 * np.abs(k2d).max() or np.abs(f - g).max() > 1e-6 * np.abs(f).max():
        return None
    return 0.5 * (f + g)


class CEMnet:
    NFFT_add = NFFT_ADD

    def __init__(self, config, upscale_kernel=None):
        self.config = config
        self.ds_factor = np.array(config.scale_factor, dtype=np.int32)
        assert np.round(self.ds_factor) == self.ds_factor, "Currently only supporting integer scale factors"
        assert upscale_kernel is None or isinstance(upscale_kernel, (str, np.ndarray)), \
            "Kernels should be given as ND-arrays, except for some specific possible strings"
        sf = int(self.ds_factor)
        self.ds_kernel = Return_kernel(sf, upscale_kernel=upscale_kernel)
        limit = config.filter_pertubation_limit
        self.ds_kernel_invalidity_half_size_LR = self._ds_margin(sf, limit)
        self.compute_inv_hTh()
        self.invalidity_margins_LR = 2 * self.ds_kernel_invalidity_half_size_LR + self.inv_hTh_invalidity_half_size
        self.invalidity_margins_HR = self.ds_factor * self.invalidity_margins_LR
        self.pre_stride, self.post_stride = sampling_phase(sf)
        # rank-1 symmetric filters (bicubic, mildly blurred bicubic) take the separable kernels; anything else
        # (estimated kernels, strong blur: the magnitude clamp of inv_hTh is not separable) the 2-D stencils
        self._ds_1d = _rank1_factor(self.ds_kernel)
        self._inv_1d = _rank1_factor(self.inv_hTh)
        self.separable = (self._ds_1d is not None and self._inv_1d is not None
                          and max(len(self._ds_1d), len(self._inv_1d)) <= capi.CEM_MAX_TAPS)
        if self.separable:
            self._filters = capi.cem_filters_struct(sf, self.pre_stride, self._ds_1d, self._inv_1d)
        else:
            self._filters = capi.CemFilterBank2D(sf, self.pre_stride, self.ds_kernel, self.inv_hTh)

    def _ds_margin(self, sf, limit):
        # downscaling a constant image with zero padding: imresize(ones, 1/sf, use_zero_padding=True)
        n = 100
        pre, _ = sampling_phase(sf)
        resp = _conv2_full(np.ones((sf * n, sf * n)), self.ds_kernel.astype(np.float64))
        o = self.ds_kernel.shape[0] // 2
        resp = resp[o:o + sf * n, o:o + sf * n][pre::sf, pre::sf]
        return _response_margin(resp, limit)

    def compute_inv_hTh(self):
        """K = (H H^T)^-1 as a spatial filter: correlate h with itself, decimate, invert in the
        Fourier domain with the magnitude clamped from below, recentre and crop."""
        sf = int(self.ds_factor)
        h = self.ds_kernel.astype(np.float64)
        hTh = _conv2_full(h, h[::-1, ::-1]) * sf ** 2
        half = int(np.ceil(hTh.shape[0] / 2))
        first = half % sf
        first = (sf if first == 0 else first) - 1
        hTh = hTh[first::sf, first::sf]
        p = self.NFFT_add // 2
        spec = np.fft.fft2(np.pad(hTh, p, mode="constant"))
        spec = spec * np.maximum(1, self.config.lower_magnitude_bound / np.abs(spec))
        inv = np.real(np.fft.ifft2(1 / spec))
        n = inv.shape[0]
        r, c = np.unravel_index(np.argmax(inv), inv.shape)
        if not (np.ceil(n / 2) == r - 1 and np.ceil(n / 2) == c - 1):
            hs = min(n - r - 1, n - c - 1, r, c)
            inv = inv[r - hs:r + hs + 1, c - hs:c + hs + 1]
        self.inv_hTh = inv
        resp = _conv2_same_ones(100, inv)
        self.inv_hTh_invalidity_half_size = _response_margin(resp, self.config.filter_pertubation_limit)
        drop = inv.shape[0] // 2 - _response_margin(resp, self.config.desired_inv_hTh_energy_portion)
        if drop > 0:
            self.inv_hTh = inv[drop:-drop, drop:-drop]

    def WrapArchitecture_PyTorch(self, generated_image=None, training_patch_size=None, only_padders=False):
        mL = int(self.invalidity_margins_LR)
        mH = int(self.ds_factor) * mL
        self.LR_padder = nn.ReplicationPad2d((mL, mL, mL, mL))
        self.HR_padder = nn.ReplicationPad2d((mH, mH, mH, mH))
        self.HR_unpadder = lambda x: x[:, :, mH:-mH, mH:-mH]
        self.LR_unpadder = lambda x: x[:, :, mL:-mL, mL:-mL]
        self.loss_mask = None
        if training_patch_size is not None:
            mask = np.zeros([1, 1, training_patch_size, training_patch_size])
            mask[:, :, mH:-mH, mH:-mH] = 1
            assert np.mean(mask) > 0, "Loss mask completely nullifies image."
            print("Using only only %.3f of patch area for learning. The rest is considered to have boundary effects"
                  % np.mean(mask))
            self.loss_mask = torch.from_numpy(mask).float()
            if torch.cuda.is_available():
                self.loss_mask = self.loss_mask.cuda()
        if only_padders:
            return None
        wrapped = CEM_PyTorch(self, generated_image)
        self.OP_names = [m[0] for m in wrapped.named_modules() if "Filter_OP" in m[0]]
        return wrapped

    def Mask_Invalid_Regions_PyTorch(self, im1, im2):
        assert self.loss_mask is not None, "Mask not defined, probably didn't pass patch size"
        return self.loss_mask * im1, self.loss_mask * im2


# --------------------------------------------------------------------------- torch side
def _require_cuda_f32(t, what):
    if not t.is_cuda:
        raise capi.EsrError("%s: expected a CUDA tensor; this package has no CPU path" % what)
    capi.require_device(t.device.index if t.device.index is not None else torch.cuda.current_device())
    return t.contiguous().float()


class Filter_Layer(nn.Module):
    """One fixed CEM filter.  ``Filter_OP.weight`` keeps the reference's [3,1,k,k] frozen parameter
    (state_dict contract); calling the layer runs the corresponding CUDA operator."""

    def __init__(self, filter, op, filters_struct):
        super().__init__()
        k = np.ascontiguousarray(np.tile(filter[None, None].astype(np.float32), (3, 1, 1, 1)))
        self.Filter_OP = nn.Conv2d(3, 3, kernel_size=filter.shape, bias=False, groups=3)
        self.Filter_OP.weight = nn.Parameter(torch.from_numpy(k), requires_grad=False)
        self.Filter_OP.filter_layer = True
        self._op, self._f = op, filters_struct

    def forward(self, x):
        if torch.is_grad_enabled() and x.requires_grad:
            raise NotImplementedError("standalone CEM operators are forward-only; gradients flow through CEM_PyTorch.forward")
        x = _require_cuda_f32(x, "CEM " + self._op)
        B, Cc, H, W = x.shape
        sf = self._f.sf
        with torch.cuda.device(x.device):
            if self._op == "down":
                out = torch.empty(B, Cc, H // sf, W // sf, device=x.device, dtype=torch.float32)
                capi.cem_call("downscale", self._f, capi.ptr(x), B, Cc, H, W, capi.ptr(out), capi.stream_ptr())
            elif self._op == "up":
                out = torch.empty(B, Cc, H * sf, W * sf, device=x.device, dtype=torch.float32)
                capi.cem_call("upscale", self._f, capi.ptr(x), B, Cc, H, W, capi.ptr(out), capi.stream_ptr())
            else:
                out = torch.empty_like(x)
                capi.cem_call("inv_hth", self._f, capi.ptr(x), B, Cc, H, W, capi.ptr(out), capi.stream_ptr())
        return out


class _CemProject(torch.autograd.Function):
    """out = crop(y + Up(K*(x - Down y))); gradient w.r.t. y only (x is the given LR image)."""

    @staticmethod
    def forward(ctx, y, x, filters, crop):
        y = _require_cuda_f32(y, "CEM projection")
        x = _require_cuda_f32(x, "CEM projection")
        B, Cc, H, W = y.shape
        sf = filters.sf
        assert x.shape == (B, Cc, H // sf, W // sf), "LR / HR size mismatch"
        out = torch.empty(B, Cc, H - 2 * crop, W - 2 * crop, device=y.device, dtype=torch.float32)
        ws = torch.empty(2 * B * Cc * (H // sf) * (W // sf), device=y.device, dtype=torch.float32)
        with torch.cuda.device(y.device):
            capi.cem_call("project", filters, capi.ptr(y), capi.ptr(x), B, Cc, H, W, crop, capi.ptr(out), capi.ptr(ws),
                          capi.stream_ptr())
        ctx.filters, ctx.crop, ctx.shape = filters, crop, (B, Cc, H, W)
        return out

    @staticmethod
    def backward(ctx, g):
        B, Cc, H, W = ctx.shape
        sf = ctx.filters.sf
        g = g.contiguous().float()
        gy = torch.empty(B, Cc, H, W, device=g.device, dtype=torch.float32)
        n = B * Cc * (H * W + H * (W // sf) + 2 * (H // sf) * (W // sf))
        ws = torch.empty(n, device=g.device, dtype=torch.float32)
        with torch.cuda.device(g.device):
            capi.cem_call("project_bwd", ctx.filters, capi.ptr(g), B, Cc, H, W, ctx.crop, capi.ptr(gy), capi.ptr(ws),
                          capi.stream_ptr())
        return gy, None, None, None


class CEM_PyTorch(nn.Module):
    def __init__(self, CEMnet, generated_image):
        super().__init__()
        self.ds_factor = CEMnet.ds_factor
        self.config = CEMnet.config
        self.generated_image_model = generated_image
        f = CEMnet._filters
        self._filters = f
        sf = int(CEMnet.ds_factor)
        self.Conv_LR_with_Inv_hTh_OP = Filter_Layer(CEMnet.inv_hTh, "inv", f)
        self.Upscale_OP = Filter_Layer(CEMnet.ds_kernel * sf ** 2, "up", f)
        self.DownscaleOP = Filter_Layer(np.rot90(CEMnet.ds_kernel, 2), "down", f)
        self.LR_padder, self.HR_padder = CEMnet.LR_padder, CEMnet.HR_padder
        self.HR_unpadder, self.LR_unpadder = CEMnet.HR_unpadder, CEMnet.LR_unpadder
        self._margin_LR = int(CEMnet.invalidity_margins_LR)
        self.pre_pad = False

    def capture(self, x_static, slot=0):
        """(graph, out) replaying forward(x_static) without autograd (inference serving, parallel.HostPipeline), or
        None when the wrapped generator is not this package's RRDBNet."""
        from .rrdbnet import RRDBNet, capture_inference
        G = self.generated_image_model
        if not isinstance(G, RRDBNet):
            return None
        return capture_inference(G, x_static, self._margin_LR if self.pre_pad else 0, self._filters, slot=slot)

    def forward(self, x):
        from .rrdbnet import RRDBNet, run_generator
        G = self.generated_image_model
        m = self._margin_LR if self.pre_pad else 0
        if isinstance(G, RRDBNet):
            return run_generator(G, x, margin=m, cem_filters=self._filters)
        # generic wrapped module: pad in the packed layout, call it, project
        x = _require_cuda_f32(x, "CEM_PyTorch.forward")
        sf = int(self.ds_factor)
        if m > 0:
            nz = 0
            if x.size(1) != 3 and x.size(1) - 3 != getattr(G, "num_latent_channels", -1):
                nz = (x.size(1) - 3) // (sf * sf)
            if x.size(1) - 3 != nz * sf * sf:
                raise NotImplementedError("LR-domain latent input with pre-padding is not built")
            B, _, h, w = x.shape
            xp = torch.empty(B, x.size(1), h + 2 * m, w + 2 * m, device=x.device, dtype=torch.float32)
            with torch.cuda.device(x.device):
                capi.check(capi.lib().esr_cem_pad_input(capi.ptr(x), B, nz, h, w, m, sf, capi.ptr(xp), capi.stream_ptr()))
            x = xp
        y = G(x)
        assert y.size(2) % sf == 0 and y.size(3) % sf == 0
        return _CemProject.apply(y, x[:, -3:].contiguous(), self._filters, sf * m)

    def train(self, mode=True):
        super().train(mode=mode)
        self.pre_pad = not mode  # pad only in eval mode (CEMnet.py:192-194)
        return self


def Adjust_State_Dict_Keys(loaded_state_dict, current_state_dict):
    """Prefix a plain-ESRGAN checkpoint's keys for the CEM-wrapped module (CEMnet.py:235-245)."""
    wrapped = all(("generated_image_model" in k or "Filter" in k) for k in current_state_dict.keys())
    if wrapped and not any("generated_image_model" in k for k in loaded_state_dict.keys()):
        out = collections.OrderedDict(("generated_image_model." + k, v) for k, v in loaded_state_dict.items())
        for k in current_state_dict.keys():
            if "Filter" in k:
                out[k] = current_state_dict[k]
        return out
    return loaded_state_dict
