"""The GUI's richer editing objectives on the device: histogram / dictionary imitation, local
STD and patch-magnitude changes, periodicity, scribbles and the "random diverse solutions" objective of the
reference's ``Z_optimization.py`` (:21-270 ``SoftHistogramLoss`` / ``ReturnPatchExtractionMat``, :371-523 the per-objective
set-up, :576-617 the per-objective loss, :525-535 ``Masked_STD``, :657-679 the sub-image helpers and ``PeriodicityLoss``).

What is different from the reference's formulation:
* the O(pixels x bins) kernel-density arithmetic of the histogram / dictionary losses is two kernels of libesr_b200.so
  (``csrc/zobj.cu``: ``esr_kde_sums`` / ``esr_kde_grad``) behind one autograd node, instead of [D, N, M] fp64 tensors;
* patch extraction is an index table + gather (``PatchTable``) instead of a sparse 0/1 matrix product, and the greedy
  patch selection is a native host loop (``esr_patch_select``);
* every objective is an object (its set-up in the constructor, ``__call__(fake_H) -> per-image loss``); ``resolve()`` maps
  the GUI's objective strings (substring conventions of the reference) onto them once.

Everything consumes the generator's output ``fake_H`` and hands back dL/d fake_H; G+CEM forward and the data gradient stay
on this package's kernels.  No CPU path: the density kernels raise on CPU tensors."""
import ctypes as C

import numpy as np
import torch

from . import _capi as capi

SQRT_EPSILON = 1e-7        # Z_optimization.py:27


# ----------------------------------------------------------------------------------------------- density sums
def _kde_sums_device(samples, bins, period, temperature, eps, per_bin):
    if not samples.is_cuda:
        raise capi.EsrError("kde_sums needs CUDA tensors: the histogram / dictionary objectives have no CPU path")
    capi.require_device(samples.device.index)
    l = capi.lib()
    D, N = samples.shape
    M = bins.shape[1]
    n_own, n_other = (M, N) if per_bin else (N, M)
    out = torch.empty(n_own, device=samples.device, dtype=torch.float64)
    ws_bytes = int(l.esr_kde_workspace_bytes(n_own, n_other))
    ws = torch.empty(max(1, ws_bytes // 8), device=samples.device, dtype=torch.float64)
    with torch.cuda.device(samples.device):
        if per_bin:
            capi.check(l.esr_kde_sums(capi.ptr(bins), 1, M, capi.ptr(samples), 0, N, D, period, temperature, eps,
                                      capi.ptr(out), capi.ptr(ws), capi.stream_ptr()))
        else:
            capi.check(l.esr_kde_sums(capi.ptr(samples), 0, N, capi.ptr(bins), 1, M, D, period, temperature, eps,
                                      capi.ptr(out), capi.ptr(ws), capi.stream_ptr()))
    return out


def _kde_grad_device(samples, bins, period, temperature, eps, w_sample, w_bin):
    l = capi.lib()
    D, N = samples.shape
    grad = torch.empty_like(samples)
    ws = torch.empty(max(1, int(l.esr_kde_grad_workspace_bytes(N, bins.shape[1], D)) // 8), device=samples.device, dtype=torch.float64)
    with torch.cuda.device(samples.device):
        capi.check(l.esr_kde_grad(capi.ptr(samples), N, capi.ptr(bins), bins.shape[1], D, period, temperature, eps,
                                  capi.ptr(w_sample), capi.ptr(w_bin), capi.ptr(grad), capi.ptr(ws), capi.stream_ptr()))
    return grad


class _KdeSums(torch.autograd.Function):
    """sum_j E[i,j] per sample (per_bin False) or sum_i E[i,j] per bin (per_bin True), fp64; differentiable in the samples."""

    @staticmethod
    def forward(ctx, samples, bins, period, temperature, eps, per_bin):
        samples = samples.detach().float().contiguous()
        ctx.save_for_backward(samples, bins)
        ctx.cfg = (float(period), float(temperature), float(eps), bool(per_bin))
        return _kde_sums_device(samples, bins, float(period), float(temperature), float(eps), bool(per_bin))

    @staticmethod
    def backward(ctx, g):
        samples, bins = ctx.saved_tensors
        period, temperature, eps, per_bin = ctx.cfg
        g = g.detach().double().contiguous()
        grad = _kde_grad_device(samples, bins, period, temperature, eps, None if per_bin else g, g if per_bin else None)
        return grad, None, None, None, None, None


def kde_sums(samples, bins, period, temperature, eps=SQRT_EPSILON, per_bin=False):
    """samples fp32 [D, N], bins fp64 [D, M] -> fp64 [M] (per_bin) or [N] sums of
    exp(-mean_d (wrapped |p - b| + eps)^2 / temperature)."""
    assert samples.dim() == 2 and bins.dim() == 2 and samples.size(0) == bins.size(0), (samples.shape, bins.shape)
    return _KdeSums.apply(samples, bins.detach().double().contiguous(), period, temperature, eps, per_bin)


# ----------------------------------------------------------------------------------------------- patch tables
class PatchTable:
    """Pixel indexes of the selected patches, [patch_size^2, n_patches] (value-major): ``extract`` is the gather that the
    reference expresses as a sparse matrix product (Z_optimization.py:266-270 builds rows d * n + p <-> pixel idx[p, d])."""

    def __init__(self, indexes, device):
        self.indexes_np = np.ascontiguousarray(indexes.T) if indexes.ndim == 2 else indexes.reshape(1, -1)
        self.idx = torch.from_numpy(self.indexes_np.astype(np.int64)).to(device)
        self.values_per_patch, self.n = self.idx.shape

    def extract(self, flat_image):
        return flat_image.reshape(-1)[self.idx]


def _opened_mask(mask, patch_size):
    from scipy.ndimage import binary_opening
    return binary_opening(np.asarray(mask).astype(bool), np.ones([patch_size, patch_size], dtype=bool))


def patch_tables(mask, patch_size, device, patches_overlap=1, return_non_covered=False):
    """ReturnPatchExtractionMat (Z_optimization.py:230-264): every patch_size^2 window that lies inside the opened mask, in
    raster order; for patches_overlap < 1 the greedy thinning of esr_patch_select.  Returns the PatchTable and, when asked,
    the table of the mask pixels no kept patch covers (None when nothing was thinned, :260-261)."""
    mask = _opened_mask(mask, patch_size)
    H, W = mask.shape
    labelled = np.where(mask, 1 + np.arange(mask.size, dtype=np.int64).reshape(H, W), 0)
    windows = np.lib.stride_tricks.sliding_window_view(labelled, (patch_size, patch_size))      # a view: nothing is copied
    inside = windows.min(axis=(2, 3)) > 0                                                        # windows wholly in the mask
    indexes = np.ascontiguousarray(windows[inside].reshape(-1, patch_size ** 2) - 1)             # raster order, only those
    non_covered = None
    if patches_overlap < 1:
        if indexes.size == 0:
            raise ValueError("no %dx%d patch fits in the mask" % (patch_size, patch_size))
        lo, hi = int(indexes.min()), int(indexes.max())
        span = hi - lo
        if span <= 0:
            raise ValueError("degenerate patch set")
        valid = np.zeros(indexes.shape[0], dtype=np.uint8)
        covered = np.zeros(span, dtype=np.uint8)
        capi.check(capi.lib().esr_patch_select(indexes.ctypes.data_as(C.c_void_p), indexes.shape[0], indexes.shape[1],
                                               float(patches_overlap), lo, span, valid.ctypes.data_as(C.c_void_p),
                                               covered.ctypes.data_as(C.c_void_p)))
        pixels = np.unique(indexes)
        slots = (pixels - lo - 1) % span
        print('%.3f of desired pixels are covered by assigned patches' % covered[slots].mean())
        indexes = indexes[valid.astype(bool)]
        if return_non_covered:
            non_covered = PatchTable(pixels[covered[slots] == 0].reshape(-1, 1), device)
    table = PatchTable(indexes, device)
    return (table, non_covered) if return_non_covered else table


# ----------------------------------------------------------------------------------------------- histogram / dictionary
class SoftHistogramLoss(torch.nn.Module):
    """Same constructor and call as the reference's class (Z_optimization.py:21-228) for gray_scale images (what
    Z_optimizer builds, :502-505).  bins / min / max: centre of the first and last of `bins` histogram bins;
    patch_size > 1 or colour turn it into a kernel density over the desired image's own pixels / patches."""

    def __init__(self, bins, min, max, desired_hist_image_mask=None, desired_hist_image=None, gray_scale=True,
                 input_im_HR_mask=None, patch_size=1, automatic_temperature=False, image_Z=None, temperature=0.05,
                 dictionary_not_histogram=False, no_patch_DC=False, no_patch_STD=False):
        super().__init__()
        if automatic_temperature:
            raise NotImplementedError("automatic histogram temperature (a second-order search through G) is not built")
        if not gray_scale:
            raise NotImplementedError("colour histograms are not built (Z_optimizer only asks for gray_scale=True)")
        assert no_patch_DC or not no_patch_STD, 'Not supporting removing of only patch STD without DC'
        device = desired_hist_image[0].device if desired_hist_image is not None else input_im_HR_mask.device
        self.device = device
        self.temperature, self.period, self.patch_size = float(temperature), float(max), int(patch_size)
        self.bin_width = (max - min) / (bins - 1)
        self.dictionary_not_histogram, self.no_patch_DC, self.no_patch_STD = dictionary_not_histogram, no_patch_DC, no_patch_STD
        self.KDE = patch_size > 1
        self.num_dims = patch_size ** 2
        self.normalizer = None
        gray = [im.mean(1, keepdim=True).reshape(-1) for im in desired_hist_image] if desired_hist_image is not None else None
        if self.KDE:
            assert gray is not None, 'Not supporting color images or patch histograms for model training loss for now'
            overlap = (self.num_dims - patch_size) / self.num_dims            # all of a patch but one row / column (:56)
            desired = torch.cat([patch_tables(m, patch_size, device, overlap).extract(g)
                                 for g, m in zip(gray, desired_hist_image_mask)], 1)                   # [D, N_desired]
            desired = self._normalise_patches(desired, fit=True)
            self.bins = self._prune(desired).double()
            self.patch_table = patch_tables(input_im_HR_mask.detach().cpu().numpy(), patch_size, device, 0.5)
            self.image_mask = None
        else:
            if gray is not None and len(gray) > 1:
                print('Not supproting multiple hist image versions for non-patch histogram/dictionary. Removing extra image versions.')
            desired = gray[0].reshape(1, -1) if gray is not None else None      # the desired image's mask is not applied (:178 skips it)
            self.bins = torch.linspace(min, max, bins).reshape(1, -1).double().to(device)
            self.image_mask = input_im_HR_mask.reshape(-1).to(device) > 0 if input_im_HR_mask is not None else None
        self.desired_hists_list = None
        if not dictionary_not_histogram and desired is not None:
            with torch.no_grad():
                self.desired_hists_list = [self._histogram(desired.float(), fit_normalizer=True, log=False)]

    # -- helpers
    def _normalise_patches(self, patches, fit=False):
        """DC / STD removal of patches [D, N] (:60-66 for the desired patches, :177-180 for the image's)."""
        if not self.no_patch_DC:
            return patches
        patches = patches - patches.mean(dim=0, keepdim=True)
        if self.no_patch_STD:
            std = torch.clamp(patches.std(dim=0, keepdim=True), min=1 / 255)
            if fit:
                self.mean_patches_STD = std.mean().item()
            patches = patches / std * self.mean_patches_STD
        return patches

    def _prune(self, desired):
        """Drops a desired sample when a LATER one lies within half a bin width in every value (Desired_Im_2_Bins,
        :106-130, with its pre-bool-dtype mask semantics).  Row blocks bound the memory of the pairwise comparison; the
        whole set is always pruned against itself (the reference splits the set into more and more sub-images until its
        [D, N, N] tensor fits, :108-126, so for large N its result depends on the free memory of the moment)."""
        D, N = desired.shape
        keep = torch.ones(N, dtype=torch.bool, device=desired.device)
        rows = builtins_max(1, (1 << 26) // builtins_max(1, N * D))       # <= 256 MiB of fp32 differences per block
        col = torch.arange(N, device=desired.device)
        for r0 in range(0, N, rows):
            blk = desired[:, r0:r0 + rows]
            close = ((blk.unsqueeze(2) - desired.unsqueeze(1)).abs() < self.bin_width / 2).all(0)        # [rows, N]
            later = col.unsqueeze(0) > torch.arange(r0, r0 + blk.size(1), device=desired.device).unsqueeze(1)
            keep[r0:r0 + blk.size(1)] = ~(close & later).any(1)
        return desired[:, keep]

    def _samples(self, image):
        """One gray image [1, H, W] -> the [D, N] samples the density runs on (:174-184)."""
        if self.patch_size > 1:
            return self._normalise_patches(self.patch_table.extract(image))
        flat = image.reshape(1, -1)
        return flat[:, self.image_mask] if self.image_mask is not None else flat

    def _histogram(self, samples, fit_normalizer, log):
        """Soft histogram of samples [D, N] over the bins, with the extra "everything else" bin of the density form
        (:195-204); normalizer: refitted for plain histograms and when asked (the desired image), else the stored one."""
        N = samples.size(1)
        hist = kde_sums(samples, self.bins, self.period, self.temperature, per_bin=True) / N
        if fit_normalizer or not self.KDE:
            self.normalizer = hist.sum() / N
        hist = (hist / self.normalizer / N).float()
        if self.KDE:
            hist = torch.cat([hist, (1 - torch.clamp(hist.sum(), max=1.0)).reshape(1)])
        return (torch.log(hist + torch.finfo(hist.dtype).eps) if log else hist).reshape(1, -1)

    def _dictionary(self, samples):
        """-log of the mean density of every sample under the dictionary atoms (:193-194) -> [1, N]."""
        return -torch.log(kde_sums(samples, self.bins, self.period, self.temperature, per_bin=False) / self.bins.size(1)).reshape(1, -1)

    def Feed_Desired_Hist_Im(self, desired_hist_image):
        with torch.no_grad():
            self.desired_hists_list = [self._histogram(im.mean(0, keepdim=True).reshape(1, -1).float(), fit_normalizer=True, log=False)
                                       for im in desired_hist_image]

    def forward(self, cur_images):
        per_image = []
        for image in cur_images:
            samples = self._samples(image.mean(0, keepdim=True))
            per_image.append(self._dictionary(samples) if self.dictionary_not_histogram else self._histogram(samples, False, True))
        if self.dictionary_not_histogram:
            return torch.cat(per_image, 0).mean(1).float()
        return torch.nn.functional.kl_div(torch.cat(per_image, 0), torch.cat(self.desired_hists_list, 0), reduction='mean').float()


import builtins as _builtins                       # the constructor's `min` / `max` arguments shadow the builtins
builtins_max = _builtins.max


# ----------------------------------------------------------------------------------------------- image helpers
def shifted_pair(image, shift):
    """(image translated by +shift, by -shift) cropped to their common support (Return_Translated_SubImage, :657-659)."""
    dy, dx = int(shift[0]), int(shift[1])
    if dy != shift[0] or dx != shift[1]:
        raise ValueError("integer periodicity needs integer shifts, got %r (use the nonInt objective)" % (shift,))

    def cut(t, sy, sx):
        H, W = t.shape[-2:]
        return t[..., (sy if sy > 0 else 0):(H + sy if sy < 0 else H), (sx if sx > 0 else 0):(W + sx if sx < 0 else W)]
    return cut(image, dy, dx), cut(image, -dy, -dx)


def masked_shift_difference(image, mask, shift):
    """mean over (C, H, W) of mask(+s) * mask(-s) * |image(+s) - image(-s)| per image (:678-679, :411-412)."""
    a, b = shifted_pair(image, shift)
    ma, mb = shifted_pair(mask, shift)
    return (ma * mb * (a - b).abs()).mean(dim=(1, 2, 3))


def rgb2hsv(rgb):
    """[H, W, 3] float -> HSV with H, S in [0, 1] and V in the input's range (the standard hexcone model, as skimage's)."""
    rgb = np.asarray(rgb, dtype=np.float64)
    v = rgb.max(-1)
    delta = v - rgb.min(-1)
    with np.errstate(invalid='ignore', divide='ignore'):
        s = np.where(delta == 0, 0.0, delta / v)
        r, g, b = rgb[..., 0], rgb[..., 1], rgb[..., 2]
        h = np.where(v == r, (g - b) / delta, np.where(v == g, 2.0 + (b - r) / delta, 4.0 + (r - g) / delta))
    h = np.where(delta == 0, 0.0, (h / 6.0) % 1.0)
    return np.stack([h, np.nan_to_num(s), v], -1)


def hsv2rgb(hsv):
    hsv = np.asarray(hsv, dtype=np.float64)
    h, s, v = hsv[..., 0], hsv[..., 1], hsv[..., 2]
    hi = np.floor(h * 6)
    f = h * 6 - hi
    p, q, t = v * (1 - s), v * (1 - f * s), v * (1 - (1 - f) * s)
    hi = hi.astype(np.int64) % 6
    sel = [np.stack(c, -1) for c in ((v, t, p), (q, v, p), (p, v, t), (p, q, v), (t, p, v), (v, p, q))]
    return np.choose(hi[..., None], sel)


# ----------------------------------------------------------------------------------------------- objectives
class LocalStd:
    """Masked_STD (:525-535): STD of fake_H * mask per image [1, B], or - 'local' objectives - the STD of every 7x7
    patch of the gray image (plus one entry for the pixels no patch covers) as [n, B]."""

    def __init__(self, zopt, local, image_mask_np, patch_size):
        self.zopt, self.local = zopt, local
        if local:
            overlap = 1 if 'STD' in zopt.objective else 0.5
            self.table, self.rest = patch_tables(image_mask_np, patch_size, zopt.device, overlap, return_non_covered=True)

    def __call__(self, first_image_only=False):
        fake_H = self.zopt.model.fake_H
        if not self.local:
            return torch.std(fake_H * self.zopt.image_mask, dim=(1, 2, 3)).view(1, -1)
        cols = []
        for im in fake_H[:1] if first_image_only else fake_H:
            gray = im.mean(dim=0)
            stds = self.table.extract(gray).std(dim=0)
            if self.rest is not None:
                stds = torch.cat([stds, self.rest.extract(gray).reshape(-1).std(dim=0, keepdim=True)], 0)
            cols.append(stds)
        return torch.stack(cols, 1)


def _std_preserving(zopt, weight, target):
    return (weight * (zopt.Masked_STD(first_image_only=False) - target) ** 2).mean(0)


class HistObjective:                                          # :476-505, :591-594
    def __init__(self, zopt, data, auto_temperature):
        o = zopt.objective
        if auto_temperature:
            raise NotImplementedError("auto_set_hist_temperature is not built")
        self.zopt, self.weight, self.keep_std = zopt, 1e4, 'localSTD' in o
        zopt.STD_PRESERVING_WEIGHT = self.weight
        zopt.loss = SoftHistogramLoss(bins=256, min=0, max=1, desired_hist_image=data['HR'] if data is not None else None,
                                      desired_hist_image_mask=data['Desired_Im_Mask'] if data is not None else None,
                                      input_im_HR_mask=zopt.image_mask, gray_scale=True, patch_size=6 if 'patch' in o else 1,
                                      temperature=5e-4 if 'hist' in o else 1e-3, dictionary_not_histogram='dict' in o,
                                      no_patch_DC='noDC' in o, no_patch_STD='no_localSTD' in o)

    def __call__(self, fake_H):
        loss = self.zopt.loss(fake_H)
        return loss + _std_preserving(self.zopt, self.weight, self.zopt.initial_STD) if self.keep_std else loss


class StdObjective:                                           # :426-435, :603-607 (global and local)
    def __init__(self, zopt, data):
        self.zopt, self.relative = zopt, any(p in zopt.objective for p in ('increase', 'decrease'))
        assert zopt.objective.replace('local_', '') in ['max_STD', 'min_STD', 'STD_increase', 'STD_decrease']
        if self.relative:
            up, inc = 'increase' in zopt.objective, data['STD_increment']
            zopt.desired_STD = zopt.initial_STD
            if inc is None:
                zopt.desired_STD *= 1.05 if up else 1 / 1.05
            else:
                zopt.desired_STD += inc if up else -inc

    def __call__(self, fake_H):
        std = self.zopt.Masked_STD(first_image_only=False)
        return ((std - self.zopt.desired_STD) ** 2 if self.relative else std).mean(0)


class MagObjective:                                           # 'local_Mag_*': :418-422, :608-613
    def __init__(self, zopt, data):
        self.zopt = zopt
        patches = zopt._std.table.extract(zopt.model.fake_H.mean(dim=1).reshape(-1))
        mean, std = patches.mean(dim=0, keepdim=True), torch.clamp(patches.std(dim=0, keepdim=True), min=1 / 255)
        step = data['STD_increment'] * (1 if 'increase' in zopt.objective else -1)
        zopt.desired_patches = (patches - mean) / std * (std + step) + mean

    def __call__(self, fake_H):
        t = self.zopt._std.table
        return torch.stack([((t.extract(im.mean(dim=0)) - self.zopt.desired_patches) ** 2).mean() for im in fake_H], 0)


class TvObjective:                                            # :474-475, :618-619
    def __init__(self, zopt):
        self.zopt = zopt
        zopt.STD_PRESERVING_WEIGHT = 100

    def __call__(self, fake_H):
        from .z_optimization import TV_Loss
        z = self.zopt
        return _std_preserving(z, z.STD_PRESERVING_WEIGHT, z.initial_STD) + TV_Loss(fake_H * z.image_mask)


class PeriodicityObjective:                                   # :436-471, :614-617, :664-679
    PLUS_MEANS_STD_INCREASE = True

    def __init__(self, zopt, data):
        o = zopt.objective
        self.zopt, self.non_int, self.plus = zopt, 'nonInt' in o, 'Plus' in o
        zopt.STD_PRESERVING_WEIGHT = 20
        if self.plus:
            zopt.desired_STD = zopt.initial_STD + data['STD_increment']
        if self.non_int:
            size = list(zopt.model.fake_H.shape[2:])
            like = zopt.model.fake_H
            self.points = [[self._grid(sign * np.array(p, dtype=np.float64), size).to(like) for sign in (1, -1)]
                           for p in data['periodicity_points']]
        else:
            self.points = [np.array(p) for p in data['periodicity_points']]
        zopt.periodicity_points = self.points

    @staticmethod
    def _grid(shift, size):
        """Sampling grid of the image translated by a fractional (dy, dx) and cropped to the common support, in
        grid_sample's normalised coordinates (:452-464: x then y, pixel k of n mapped to 2k/n - 1)."""
        axes = []
        for axis, s in ((0, shift[1]), (1, shift[0])):                 # x range first, then y; the reference sizes the x
            n = size[axis]                                             # range with the image HEIGHT and y with the width (:455)
            start, stop = (s if s > 0 else 0.0), (n + s if s < 0 else float(n))
            count = n - int(np.ceil(np.abs(np.array([0, n]) - np.array([start, stop]))).astype(np.int16).max())
            axes.append(np.linspace(start, stop, num=count) / n * 2 - 1)
        gx, gy = np.meshgrid(*axes)
        return torch.from_numpy(np.stack([gx, gy], -1)).unsqueeze(0)

    def __call__(self, fake_H):
        z = self.zopt
        loss = 0 if self.plus else (z.STD_PRESERVING_WEIGHT * (z.Masked_STD(first_image_only=False) - z.initial_STD) ** 2).mean()
        mask = z.image_mask.unsqueeze(0).unsqueeze(0)
        for point in self.points:
            if self.non_int:
                sample = lambda t, g: torch.nn.functional.grid_sample(t, g.repeat([t.size(0), 1, 1, 1]))      # noqa: E731
                both = sample(mask, point[0]) * sample(mask, point[1])
                loss = loss + (both * (sample(fake_H, point[0]) - sample(fake_H, point[1])).abs()).mean(dim=(1, 2, 3))
            else:
                loss = loss + masked_shift_difference(fake_H, mask, point)
        if self.plus:
            loss = loss + z.STD_PRESERVING_WEIGHT * ((z.Masked_STD(first_image_only=False) - z.desired_STD) ** 2).mean()
        return loss


class ScribbleObjective:                                      # :371-416; mask ids: 1 colour, 2 / 3 brighter / darker, > 3 smooth
    HALF_NEIGHBOURS = ((-1, -1), (-1, 0), (0, -1), (1, -1))   # 4 of the 8 neighbour differences, each pair once (:404-409)

    def __init__(self, zopt, data):
        from scipy.signal import convolve2d
        self.zopt = zopt
        dev, mask = zopt.device, zopt.image_mask
        ids_np = np.asarray(data['scribble_mask'])
        ids = torch.from_numpy(ids_np).to(mask.dtype).to(dev)
        gain = np.ones_like(ids_np, dtype=np.float32) + data['brightness_factor'] * ((ids_np == 2).astype(np.float32) - (ids_np == 3))
        gain = convolve2d(np.pad(gain, 1, mode='edge'), np.full([3, 3], 1 / 9), mode='valid')
        self.l1_mask = mask * ((ids > 0) & (ids < 4)).float()
        self.tv_masks = [(mask * (ids == i).float()).unsqueeze(0).unsqueeze(0) for i in torch.unique(ids * mask) if i > 3]
        hsv = rgb2hsv(np.clip(255 * zopt.model.fake_H[0].detach().cpu().numpy().transpose(1, 2, 0), 0, 255))
        hsv[:, :, 2] *= gain
        relit = torch.from_numpy(hsv2rgb(hsv).transpose(2, 0, 1)[None] / 255).to(mask.dtype).to(dev)
        relight = ((ids == 2) | (ids == 3)).float()
        zopt.GT_HR = zopt.GT_HR.to(dev) * (1 - relight) + relight * relit

    def __call__(self, fake_H):
        target = self.zopt.GT_HR * self.l1_mask
        rows = []
        for i in range(fake_H.size(0)):
            im = fake_H[i:i + 1]
            value = torch.nn.functional.l1_loss(im * self.l1_mask, target)
            if self.tv_masks:
                value = value + sum(masked_shift_difference(im, m, s) for m in self.tv_masks for s in self.HALF_NEIGHBOURS)
            rows.append(value)
        return torch.stack(rows, 0)


class DiverseObjective:                                       # 'random_l1' / 'random_l1_limited': :578-590
    def __init__(self, zopt, data):
        self.zopt, self.limited = zopt, 'limited' in zopt.objective
        if 'VGG' in zopt.objective:
            raise NotImplementedError("VGG feature distances are outside the built path (define_F is broken in the reference too)")
        if self.limited:
            zopt.initial_image = 1 * zopt.model.fake_H.detach()
            zopt.rmse_weight = data['rmse_weight']

    def __call__(self, fake_H):
        z, B = self.zopt, fake_H.size(0)
        apart = (fake_H.unsqueeze(0) - fake_H.unsqueeze(1)).abs() + torch.eye(B, device=fake_H.device).view(B, B, 1, 1, 1)
        loss = apart.min(dim=0)[0]
        if self.limited:
            loss = loss - z.rmse_weight * (fake_H - z.initial_image).abs()
        if z.Z_mask is not None:
            loss = loss * z.Z_mask
        return -1 * loss.mean(dim=(1, 2, 3))


def unsupported_reason(objective, auto_temperature=False):
    """None when resolve() can build `objective`, else why not (one line)."""
    o = objective
    for phrase, why in (('desired_SVD', 'the structure-tensor FilterLoss'), ('Adversarial', 'the critic objective'),
                        ('VGG', 'the VGG feature extractor (broken in the reference too)')):
        if phrase in o:
            return "needs %s, which is outside the built path" % why
    if 'random' in o and 'l1' not in o:
        return "only the l1 form of the diverse-solutions objective is built"
    if ('hist' in o or 'dict' in o) and auto_temperature:
        return "auto_set_hist_temperature (a second-order search through G) is not built"
    if not any(p in o for p in ('random', 'scribble', 'l1', 'hist', 'dict', 'STD', 'Mag', 'periodicity', 'TV')):
        return "is not one of the reference's objectives"
    return None


def resolve(zopt, data, auto_temperature=False):
    """The objective object for zopt.objective, following the precedence of the reference's substring tests (:576-619 for
    the loss, :368-509 for the set-up).  NotImplementedError names what is not built."""
    o = zopt.objective
    why = unsupported_reason(o, auto_temperature)
    if why is not None:
        raise NotImplementedError("Z objective %r %s" % (o, why))
    if 'random' in o:
        return DiverseObjective(zopt, data)
    if 'scribble' in o:
        return ScribbleObjective(zopt, data)
    if 'hist' in o or 'dict' in o:
        return HistObjective(zopt, data, auto_temperature)
    if 'STD' in o and not any(p in o for p in ('periodicity', 'TV')):
        return StdObjective(zopt, data)
    if 'Mag' in o:
        return MagObjective(zopt, data)
    if 'periodicity' in o:
        return PeriodicityObjective(zopt, data)
    if 'TV' in o:
        return TvObjective(zopt)
    raise NotImplementedError("Z objective %r is not one of the reference's objectives" % o)
