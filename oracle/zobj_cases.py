"""Seeded inputs and case tables of the rich-objective parity tests (TEST INFRASTRUCTURE; shared by oracle/gen_golden.py,
which runs the unmodified reference on them, and by tests/test_zobjectives.py / tests/test_gpu_zobjectives.py, which run
the product on the same inputs).  Nothing here is imported by the product."""
import numpy as np
import torch


def smooth_image(seed, B, H, W, amp=0.25):
    """Low-frequency images in [0, 1]: neighbouring pixels and patches resemble each other, so the density kernels of the
    histogram / dictionary objectives (temperatures 5e-4 / 1e-3) work in their non-underflowing regime."""
    rng = np.random.default_rng(seed)
    coarse = torch.from_numpy(rng.random((B, 3, max(2, H // 6), max(2, W // 6))).astype(np.float32))
    img = torch.nn.functional.interpolate(coarse, size=(H, W), mode="bicubic", align_corners=False)
    img = 0.5 + amp * (img - 0.5) + 0.01 * torch.from_numpy(rng.standard_normal((B, 3, H, W)).astype(np.float32))
    return img.clamp(0, 1).contiguous()


def hist_masks(H, W):
    """(desired-image mask (numpy bool), image mask (float tensor)) with holes at the borders."""
    dm = np.ones((H, W), dtype=bool)
    dm[:3, :] = False
    dm[:, -2:] = False
    im = torch.ones(H, W)
    im[:, :4] = 0
    im[-3:, :] = 0
    return dm, im


# name -> SoftHistogramLoss keyword arguments (Z_optimization.py:502-505 builds patch_size 6 / 1, temperatures 5e-4 / 1e-3)
HIST_CASES = {
    "hist": dict(patch_size=1, temperature=5e-4),
    "dict": dict(patch_size=1, temperature=1e-3, dictionary_not_histogram=True),
    "patchhist": dict(patch_size=6, temperature=5e-4),
    "patchhist_noDC": dict(patch_size=6, temperature=5e-4, no_patch_DC=True),
    "patchdict_noDC": dict(patch_size=6, temperature=1e-3, dictionary_not_histogram=True, no_patch_DC=True),
    "patchhist_noDC_noSTD": dict(patch_size=6, temperature=5e-4, no_patch_DC=True, no_patch_STD=True),
    "patchdict_noDC_noSTD": dict(patch_size=6, temperature=1e-3, dictionary_not_histogram=True, no_patch_DC=True, no_patch_STD=True),
}
HIST_HW = (26, 30)


def hist_inputs(name):
    """(images [2,3,H,W], desired image versions (list of [1,3,H,W]), desired masks, image mask) of a HIST_CASES entry."""
    H, W = HIST_HW
    seed = sorted(HIST_CASES).index(name)
    img = smooth_image(100 + seed, 2, H, W)
    desired = [(smooth_image(100 + seed, 1, H, W) + 0.02 * smooth_image(200 + seed + k, 1, H, W)).clamp(0, 1) for k in range(2)]
    dm, im = hist_masks(H, W)
    return img, desired, [dm, dm.copy()], im


class StubGenerator(torch.nn.Module):
    """A cheap differentiable stand-in for G+CEM on CPU (model_input [B, 48 + 3, h, w] -> [B, 3, 4h, 4w]): nearest x4 of
    the LR image plus a fixed 3x3 convolution of the un-viewed Z.  Lets the host logic of the objectives run against the
    reference's Z_optimizer without the generator's CPU cost; the GPU tests use the real generator."""

    def __init__(self):
        super().__init__()
        rng = np.random.default_rng(7)
        self.conv = torch.nn.Conv2d(3, 3, 3, padding=1)
        self.conv.weight.data = torch.from_numpy(rng.standard_normal((3, 3, 3, 3)).astype(np.float32)) * 0.08
        self.conv.bias.data.zero_()

    def forward(self, x):
        B, _, h, w = x.shape
        z = x[:, :-3].reshape(B, 3, 4 * h, 4 * w)
        lr = torch.nn.functional.interpolate(x[:, -3:], scale_factor=4, mode="nearest")
        return lr + self.conv(z)


def region_masks(H, W):
    """image mask / Z mask (numpy float32, HR size) of the masked objectives: a box well inside the image."""
    im = np.zeros((H, W), dtype=np.float32)
    im[H // 8:7 * H // 8, W // 6:5 * W // 6] = 1
    zm = np.zeros((H, W), dtype=np.float32)
    zm[H // 16:15 * H // 16, W // 12:11 * W // 12] = 1
    return im, zm


def scribble_mask(H, W):
    """ids: 0 untouched, 1 colour scribble, 2 brighter, 3 darker, 4 / 5 two smoothing (TV) regions."""
    s = np.zeros((H, W), dtype=np.float32)
    s[H // 4:H // 4 + 5, W // 4:W // 2] = 1
    s[H // 2:H // 2 + 4, W // 4:W // 4 + 8] = 2
    s[H // 2:H // 2 + 4, W // 2:W // 2 + 8] = 3
    s[5 * H // 8:5 * H // 8 + 5, W // 4:W // 4 + 6] = 4
    s[5 * H // 8:5 * H // 8 + 5, W // 2:W // 2 + 6] = 5
    return s


# name: (objective string as GUI.py:1505-1517 composes it, max_iters, batch size, extra data keys)
ZOPT3_CASES = {
    "local_std_up": ("local_STD_increase", 3, 1, {"STD_increment": 0.02}),
    "local_std_max": ("max_local_STD", 3, 1, {}),
    "local_mag_up": ("local_Mag_increase", 3, 1, {"STD_increment": 0.02}),
    "local_tv": ("local_STD_TV", 3, 1, {}),
    "periodicity_int": ("periodicity", 3, 1, {"periodicity_points": [[3, 2], [0, 4]]}),
    "periodicity_nonint": ("nonInt_periodicity", 3, 1, {"periodicity_points": [[2.5, 1.25], [-1.5, 3.0]]}),
    "periodicity_plus": ("local_STD_nonInt_periodicityPlus", 3, 1, {"periodicity_points": [[2.5, 1.25]], "STD_increment": 0.02}),
    "scribble": ("scribble", 3, 1, {"brightness_factor": 0.2}),
    "random_l1": ("random_l1", 3, 3, {}),
    "random_l1_limited": ("random_l1_limited", 3, 3, {"rmse_weight": 0.5}),
    # plain 'hist' / 'patchhist' return a 0-d KL divergence, which the reference's loop cannot iterate over (:622); the
    # forms with the STD-preserving term and the dictionary forms are per-image and run there
    "hist_keepstd": ("hist_localSTD", 3, 1, {}),
    "patchhist_noDC_keepstd": ("patchhist_noDC_no_localSTD", 3, 1, {}),
    "dict": ("dict", 3, 1, {}),
    "patchdict_noDC": ("patchdict_noDC", 3, 1, {}),
    "patchdict_noDC_keepstd": ("patchdict_noDC_no_localSTD", 3, 1, {}),
}
# the reference's scribble set-up subtracts a bool mask (`1 - desired_RGB_mask`, :396), which torch >= 1.2 rejects, and it
# needs skimage: no reference run exists for it here
NO_REFERENCE_RUN = ("scribble",)


def zopt3_data(name, lr, fake_H0, H, W):
    """The `data` dict additions and constructor keywords of a ZOPT3 case; fake_H0 is the initial output [B, 3, H, W]."""
    objective, max_iters, bs, extra = ZOPT3_CASES[name]
    data = dict(extra)
    im, zm = region_masks(H, W)
    if "scribble" in objective:
        data["scribble_mask"] = scribble_mask(H, W)
        data["HR"] = (0.8 * fake_H0[:1].detach() + 0.1).clamp(0, 1)
    if "hist" in objective or "dict" in objective:
        seed = sorted(ZOPT3_CASES).index(name)
        data["HR"] = [(fake_H0[:1].detach() + 0.03 * (smooth_image(300 + seed + k, 1, H, W) - 0.5)).clamp(0, 1) for k in range(2)]
        dm = np.ones((H, W), dtype=bool)
        dm[:2] = False
        data["Desired_Im_Mask"] = [dm, dm.copy()]
    return data, dict(image_mask=im, Z_mask=zm)


def run_hist_case(loss_class, name, device="cpu"):
    """Builds `loss_class` (the reference's or the product's SoftHistogramLoss) on a HIST_CASES entry and returns
    (per-call value, d mean(value) / d images, the bins as [D, M])."""
    img, desired, dmasks, im = hist_inputs(name)
    x = img.to(device).requires_grad_(True)
    loss = loss_class(bins=256, min=0, max=1, desired_hist_image=[d.to(device) for d in desired],
                      desired_hist_image_mask=[m.copy() for m in dmasks], input_im_HR_mask=im.to(device), gray_scale=True,
                      **HIST_CASES[name])
    value = loss(x)
    grad, = torch.autograd.grad(value.mean(), x)
    bins = loss.bins.reshape(loss.bins.shape[0], -1)
    return value.detach().cpu(), grad.cpu(), bins.detach().cpu()


ZOPT3_HW = (8, 8)            # LR size; HR 32 x 32


def run_zopt_case(optimizer_class, model, netG, name, lr, z0, device="cpu", z_init=None):
    """One ZOPT3 case through `optimizer_class` (the reference's or the product's Z_optimizer) around `model` / `netG`.
    Returns the optimiser after optimize() and the Z it returned.  z_init: per-image starting Z for batch sizes > 1
    (replaces the RNG draw of Randomize_Z, Z_optimization.py:559-560)."""
    objective, max_iters, bs, _ = ZOPT3_CASES[name]
    H, W = 4 * lr.shape[2], 4 * lr.shape[3]
    data = {"LR": lr.repeat(bs, 1, 1, 1).to(device), "Z": (0.5 * z0).repeat(bs, 1, 1, 1).to(device)}
    model.feed_data(data)
    with torch.no_grad():
        model.fake_H = netG(model.model_input)
    extra, kw = zopt3_data(name, lr, model.fake_H.cpu(), H, W)
    for k, v in extra.items():
        data[k] = v.to(device) if torch.is_tensor(v) else ([t.to(device) for t in v] if isinstance(v, list) and torch.is_tensor(v[0]) else v)
    torch.manual_seed(3)
    opt = optimizer_class(objective=objective, Z_size=[H, W], model=model, Z_range=1.0, max_iters=max_iters, data=data,
                          initial_LR=0.1, batch_size=bs, initial_Z=(0.5 * z0).to(device), **kw)
    if bs > 1:
        opt.random_Z_inits = False
        opt.Z_model.Z.data.copy_(z_init.to(device))
    Z = opt.optimize()
    return opt, Z


def zopt3_z_init(name):
    bs = ZOPT3_CASES[name][2]
    H, W = 4 * ZOPT3_HW[0], 4 * ZOPT3_HW[1]
    return torch.from_numpy(np.random.default_rng(23).standard_normal((bs, 3, H, W)).astype(np.float32)) * 0.3
