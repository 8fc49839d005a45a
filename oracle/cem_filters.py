"""Oracle: derivation of the three fixed CEM filters (float64 numpy).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates, for the default bicubic
kernel and integer scale factors:

* ``Cubic_Kernel``            codes/CEM/imresize_CEM.py:88-94   (cv2 INTER_CUBIC of a
  delta image == Keys cubic, a=-0.75, sampled at (i+0.5)/sf-0.5; cv2 4.x
  ``interpolateCubic``; cv2 is un-vendored, so its published formula is restated)
* ``calc_strides``            codes/CEM/imresize_CEM.py:73-86
* ``imresize(..., return_upscale_kernel=True)``  codes/CEM/imresize_CEM.py:18-47
* ``Return_kernel``           codes/CEM/CEMnet.py:218-219
* ``compute_inv_hTh``         codes/CEM/CEMnet.py:105-126
* ``Return_Invalid_Margin_Size_in_LR``  codes/CEM/CEMnet.py:28-42
* margins                     codes/CEM/CEMnet.py:23-26
"""
import numpy as np
from scipy.signal import convolve2d

NFFT_ADD = 36  # CEMnet.NFFT_add, codes/CEM/CEMnet.py:15


def keys_cubic(d, a=-0.75):
    d = np.abs(np.asarray(d, dtype=np.float64))
    near = ((a + 2.0) * d - (a + 3.0)) * d * d + 1.0
    far = ((a * d - 5.0 * a) * d + 8.0 * a) * d - 4.0 * a
    return np.where(d <= 1.0, near, np.where(d < 2.0, far, 0.0))


def cubic_upscale_kernel(sf):
    """imresize_CEM.py:88-94.  2-D upscale kernel (sums to sf**2)."""
    delta_size = 11
    centre = int(np.ceil(delta_size / 2)) - 1
    dst = np.arange(sf * delta_size)
    w = keys_cubic((dst + 0.5) / sf - 0.5 - centre)
    support = np.nonzero(w)[0]
    w = w[support[0]:support[-1] + 1]
    return np.outer(w, w)


def calc_strides(sf):
    """imresize_CEM.py:73-86, align_center=False branch."""
    post = int(np.floor(sf / 2))
    pre = sf - post - 1
    return pre, post


def upscale_antialiasing_kernel(sf):
    """imresize_CEM.py:18-47 with return_upscale_kernel=True and scale_factor=sf>1."""
    pre, post = calc_strides(sf)
    k = cubic_upscale_kernel(sf)
    post_pad, pre_pad = max(0, pre - post), max(0, post - pre)
    return np.pad(k, ((pre_pad, post_pad), (pre_pad, post_pad)), mode="constant")


def ds_kernel(sf):
    """CEMnet.py:218-219."""
    return (np.rot90(upscale_antialiasing_kernel(sf), 2).astype(np.float32) / (sf ** 2))


def _downscale_zero_pad_ones(sf, n):
    """imresize(ones, [1/sf], use_zero_padding=True): imresize_CEM.py:43-70."""
    pre, _ = calc_strides(sf)
    k = np.rot90(upscale_antialiasing_kernel(sf) * (1.0 / sf) ** 2, 2)
    out = convolve2d(np.ones([sf * n, sf * n]), k, mode="same")
    return out[pre::sf, pre::sf]


def invalid_margin(response, limit):
    """CEMnet.py:28-42 applied to the response of a filter to an all-ones image."""
    n = response.shape[0]
    r = response / response[n // 2, n // 2]
    r = np.where(r <= 0, limit / 2, r)
    bad = np.exp(-np.abs(np.log(r))) < limit
    rows = np.argwhere(bad[:n // 2, n // 2])
    cols = np.argwhere(bad[n // 2, :n // 2])
    return int(max(rows[-1][0] + 1, cols[-1][0] + 1))


def aliased_down_sampling(arr, sf):
    """CEMnet.py:196-203 with calc_strides(..., align_center=True)."""
    half = np.ceil(np.array(arr.shape[:2]) / 2)
    pre = np.mod(half, sf)
    pre[pre == 0] = sf
    pre = (pre - 1).astype(np.int32)
    return arr[pre[0]::sf, pre[1]::sf]


def derive(sf, lower_magnitude_bound=0.01, perturbation_limit=0.999, energy_portion=1 - 1e-6):
    """Returns dict(ds_kernel, inv_hTh, margin_LR, margin_HR, ds_half, inv_half)."""
    test = 100
    h = ds_kernel(sf)
    ds_half = invalid_margin(_downscale_zero_pad_ones(sf, test), perturbation_limit)
    hTh = convolve2d(h, np.rot90(h, 2)) * sf ** 2
    hTh = aliased_down_sampling(hTh, sf)
    p = NFFT_ADD // 2
    f = np.fft.fft2(np.pad(hTh, ((p, p), (p, p)), mode="constant"))
    f = f * np.maximum(1, lower_magnitude_bound / np.abs(f))
    inv = np.real(np.fft.ifft2(1 / f))
    n = inv.shape[0]
    mr, mc = np.argmax(inv) // n, np.mod(np.argmax(inv), n)
    if not np.all(np.equal(np.ceil(np.array(inv.shape) / 2), np.array([mr, mc]) - 1)):
        hs = min(n - mr - 1, n - mc - 1, mr, mc)
        inv = inv[mr - hs:mr + hs + 1, mc - hs:mc + hs + 1]
    ones = np.ones([test, test])
    inv_half = invalid_margin(convolve2d(ones, inv, mode="same"), perturbation_limit)
    drop = inv.shape[0] // 2 - invalid_margin(convolve2d(ones, inv, mode="same"), energy_portion)
    if drop > 0:
        inv = inv[drop:-drop, drop:-drop]
    margin_lr = 2 * ds_half + inv_half
    return dict(ds_kernel=h, inv_hTh=inv, margin_LR=int(margin_lr), margin_HR=int(sf * margin_lr),
                ds_half=ds_half, inv_half=inv_half)
