"""Import the UNMODIFIED reference (read-only, /root/reference/codes) on a CPU-only box.

TEST INFRASTRUCTURE, build-container only: /root/reference does not exist on the
GPU box, so nothing in the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` calls this
module with /root/reference (the vendored copy of oracle/vendor_ref.py is what the GPU box sees).  It is used by ``oracle/gen_golden.py`` to produce the committed fixtures
and by ``tests/test_checkpoint.py::test_matches_reference_loader`` (skipped when the reference is absent).

The shims only repair imports that broke with newer library versions or that
assume a CUDA device; no reference source is modified or copied.
"""
import os
import sys
import types

_VENDORED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "codes")   # oracle/vendor_ref.py (GPU box)
REF_ROOT = os.environ.get("ESR_REFERENCE_ROOT") or \
    ("/root/reference/codes" if os.path.isdir("/root/reference/codes/CEM") else _VENDORED)
COMPAT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          "explorable-super-resolution_old_b200", "compat")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "CEM"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


def install(compat_first=False):
    """Make ``import CEM.CEMnet``, ``models.networks``, ``Z_optimization`` work.  compat_first: put this package's
    import-path shims AHEAD of the reference (the drop-in configuration of INTEGRATION.md): the reference's callers
    then run on the B200 implementations of the hot path."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    import numpy as np
    import scipy.signal
    import scipy.signal.windows
    import torch

    if not hasattr(scipy.signal, "gaussian"):  # removed in scipy >= 1.13 (imresize_CEM.py:4)
        scipy.signal.gaussian = scipy.signal.windows.gaussian
    if not hasattr(np, "bool"):  # Z_optimization.py:232
        np.bool = bool
    if not torch.cuda.is_available():  # CEMnet.py:74,134 / Z_optimization.py:14,276
        torch.cuda.FloatTensor = torch.FloatTensor
        torch.cuda.DoubleTensor = torch.DoubleTensor
    # optional third-party modules the hot path never executes
    try:
        import skimage.color  # noqa: F401
    except Exception:
        sk = _stub("skimage")
        sk.color = _stub("skimage.color", rgb2hsv=None, hsv2rgb=None)
        sk.transform = _stub("skimage.transform", resize=None)
        sk.measure = _stub("skimage.measure")
    for name in ("GPUtil", "matplotlib", "matplotlib.pyplot", "tensorboard_logger", "lmdb", "imageio"):
        try:
            __import__(name)
        except Exception:
            _stub(name)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if compat_first:
        if COMPAT_DIR in sys.path:
            sys.path.remove(COMPAT_DIR)
        sys.path.insert(sys.path.index(REF_ROOT), COMPAT_DIR)


class _CpuDeviceTorch:
    """Module-local ``torch`` proxy: torch.device('cuda') -> cpu (Z_optimization.py:29,345)."""

    def __init__(self, torch):
        self._t = torch

    def __getattr__(self, k):
        return getattr(self._t, k)

    def device(self, *a, **kw):
        if a and isinstance(a[0], str) and a[0].startswith("cuda") and not self._t.cuda.is_available():
            return self._t.device("cpu")
        return self._t.device(*a, **kw)

    def triu(self, x, *a, **kw):
        """Z_optimization.py:118-120 builds a keep-mask as `torch.triu(mat).any(1) ^ 1`, written for torch < 1.2 where
        comparisons were uint8: there the result is a uint8 MASK.  With bool tensors `^ 1` promotes to int64 and the
        next line's `im[:, mask]` silently becomes a gather of columns 0 / 1.  Handing triu's result back as uint8
        restores the dtype the line was written for (`.any()` of uint8 stays uint8), without touching the source."""
        r = self._t.triu(x, *a, **kw)
        return r.to(self._t.uint8) if r.dtype == self._t.bool else r


def load_reference():
    """Returns (CEMnet module, networks module, architecture module, Z_optimization module)."""
    install()
    import torch
    import CEM.CEMnet as CEMnet
    import models.networks as networks
    import models.modules.architecture as arch
    import Z_optimization as zopt
    zopt.torch = _CpuDeviceTorch(torch)        # device('cuda') -> cpu only when there is no GPU; the triu dtype repair always
    return CEMnet, networks, arch, zopt


def make_opt(nb, latent_input, is_train=False, patch=256, sf=4):
    return {"gpu_ids": None, "is_train": is_train, "datasets": {"train": {"patch_size": patch}},
            "network_G": dict(which_model_G="RRDB_net", CEM_arch=1, latent_input=latent_input,
                              latent_input_domain="HR_downscaled", latent_channels=3, norm_type=None, mode="CNA",
                              nf=64, nb=nb, in_nc=3, out_nc=3, gc=32, scale=sf)}


def build_ref_G(CEMnet, networks, nb, latent_input, kind, seed, sf=4):
    """The reference's define_G -> CEM_PyTorch(RRDBNet) on CPU with the synthetic weights of esr_b200.synth."""
    import CEM.imresize_CEM as im
    from esr_b200 import synth
    im.imresize.kernels = {}
    cem = CEMnet.CEMnet(CEMnet.Get_CEM_Config(sf))
    netG = networks.define_G(make_opt(nb, latent_input, sf=sf), CEM=cem, num_latent_channels=0 if latent_input == "None" else 3)
    li = None if latent_input == "None" else latent_input + "_HR_downscaled"
    w = synth.make_weights(kind, seed=seed, nb=nb, latent_input=li, upscale=sf)
    sd = netG.state_dict()
    assert [k for k in sd if "Filter" not in k] == ["generated_image_model." + k for k in w]
    sd.update({"generated_image_model." + k: v for k, v in w.items()})
    netG.load_state_dict(sd)
    return netG, cem
