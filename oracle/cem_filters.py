"""Oracle: derivation of the three fixed CEM filters (float64 numpy).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates, for integer scale factors and the
default bicubic kernel, ``blurry_cubic_<sigma>`` and user-supplied ndarray kernels:

* ``Cubic_Kernel``            codes/CEM/imresize_CEM.py:88-94   (cv2 INTER_CUBIC of a
  delta image == Keys cubic, a=-0.75, sampled at (i+0.5)/sf-0.5; cv2 4.x
  ``interpolateCubic``; cv2 is un-vendored, so its published formula is restated)
* ``calc_strides``            codes/CEM/imresize_CEM.py:73-86
* ``imresize(..., return_upscale_kernel=True)``  codes/CEM/imresize_CEM.py:18-47
* ``Return_kernel``           codes/CEM/CEMnet.py:218-219
* ``compute_inv_hTh``         codes/CEM/CEMnet.py:105-126
* ``Return_Invalid_Margin_Size_in_LR``  codes/CEM/CEMnet.py:28-42
* margins                     codes/CEM/CEMnet.py:23-26
* ``Gaussian_2D``             codes/CEM/imresize_CEM.py:101-109  (blurry_cubic, :37-41)
* ``Center_Mass``             codes/CEM/imresize_CEM.py:114-160  (ndarray kernels, :22-32)
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


def gaussian_2d(sigma):
    """imresize_CEM.py:101-109 with size=None: odd support leaving 0.5% of the 1-D mass on each side."""
    from scipy.signal.windows import gaussian
    from scipy.stats import norm
    size = int(1 + 2 * np.ceil(-1 * norm.ppf(0.005, scale=sigma)))
    g = gaussian(size, sigma)
    g2 = g.reshape([1, size]) * g.reshape([size, 1])
    return g2 / np.sum(g2)


def energy_profile(k):
    """imresize_CEM.py:162-164: root-energy left after peeling f frames off the filter, relative to all of it."""
    e = [np.sqrt(np.sum(k ** 2))]
    for f in range(1, int(np.ceil(k.shape[0] / 2))):
        e.append(np.sqrt(np.sum(k[f:-f, f:-f] ** 2)))
    return np.array(e) / e[0]


def center_mass(k, sf):
    """imresize_CEM.py:114-160: zero-pad so that the centre of mass becomes the centre of the array, keep it
    square, trim to 99% root-energy with a side compatible with the scale factor, renormalise."""
    n = k.shape[0]
    assert k.shape[0] == k.shape[1]
    xg, yg = np.meshgrid(np.arange(n), np.arange(n))
    xc = convolve2d(xg, k, mode="valid")[0, 0] + 1
    yc = convolve2d(yg, k, mode="valid")[0, 0] + 1
    x_pad, y_pad = 2 * (n / 2 - xc), 2 * (n / 2 - yc)
    diff = np.round(np.abs(y_pad)) - np.round(np.abs(x_pad))
    px = [max(0, -x_pad), max(0, x_pad)]
    py = [max(0, -y_pad), max(0, y_pad)]

    def widen(p, d):
        to_right = np.round(p[1]) - p[1] - (np.round(p[0]) - p[0])
        q = [int(np.round(p[0])), int(np.round(p[1]))]
        big, small = int(np.ceil(d / 2)), int(np.floor(d / 2))
        if to_right > 0:
            return [q[0] + small, q[1] + big]
        return [q[0] + big, q[1] + small]
    if diff > 0:
        px = widen(px, diff)
    elif diff < 0:
        py = widen(py, -diff)
    px = [int(np.round(v)) for v in px]
    py = [int(np.round(v)) for v in py]
    k = np.pad(k, (tuple(py), tuple(px)), mode="constant")
    assert k.shape[0] == k.shape[1]
    first_below = int(np.argwhere(energy_profile(k) < 0.99)[0][0])
    drop = [first_below, first_below]
    turn = 0
    while np.mod(k.shape[0] - sum(drop) - 1 + np.mod(sf + 1, 2), sf) != 0:
        drop[turn] -= 1
        turn = (turn + 1) % 2
    k = k[drop[0]:-drop[1], drop[0]:-drop[1]]
    return k / np.sum(k)


def upscale_antialiasing_kernel(sf, kernel=None):
    """imresize_CEM.py:18-47 with return_upscale_kernel=True and scale_factor=sf>1.
    kernel: None (bicubic), 'blurry_cubic_<sigma>' or a square ndarray downscaling kernel."""
    pre, post = calc_strides(sf)
    post_pad, pre_pad = max(0, pre - post), max(0, post - pre)
    if isinstance(kernel, np.ndarray):
        assert np.abs(1 - np.sum(kernel)) < np.finfo(np.float32).eps
        k = center_mass(np.rot90(kernel, 2), sf) * sf ** 2
        assert (k.shape[0] + post_pad + pre_pad - 1) % sf == 0
    else:
        k = cubic_upscale_kernel(sf)
        if kernel is not None and "blurry_cubic" in kernel:
            k = convolve2d(k, gaussian_2d(float(kernel[len("blurry_cubic_"):])))
    return np.pad(k, ((pre_pad, post_pad), (pre_pad, post_pad)), mode="constant")


def ds_kernel(sf, kernel=None):
    """CEMnet.py:218-219."""
    return (np.rot90(upscale_antialiasing_kernel(sf, kernel), 2).astype(np.float32) / (sf ** 2))


def _downscale_zero_pad_ones(sf, n, kernel=None):
    """imresize(ones, [1/sf], use_zero_padding=True): imresize_CEM.py:43-70."""
    pre, _ = calc_strides(sf)
    k = np.rot90(upscale_antialiasing_kernel(sf, kernel) * (1.0 / sf) ** 2, 2)
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


def derive(sf, lower_magnitude_bound=0.01, perturbation_limit=0.999, energy_portion=1 - 1e-6, kernel=None):
    """Returns dict(ds_kernel, inv_hTh, margin_LR, margin_HR, ds_half, inv_half)."""
    test = 100
    h = ds_kernel(sf, kernel)
    ds_half = invalid_margin(_downscale_zero_pad_ones(sf, test, kernel), perturbation_limit)
    # the reference divides the float32 kernel by a 0-d int32 array, which numpy >= 2 promotes to float64:
    # everything downstream runs in float64 on float32-rounded taps (the energy-portion crop is sensitive to it)
    h64 = h.astype(np.float64)
    hTh = convolve2d(h64, np.rot90(h64, 2)) * sf ** 2
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
