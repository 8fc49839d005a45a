"""`from Z_optimization import Z_optimizer`.

``esr_b200.z_optimization.Z_optimizer`` runs every objective this package builds: l1, TV and the STD objectives (global
ones as one CUDA graph per iteration), and through ``esr_b200.z_objectives`` the local STD / Mag variants, histogram and
dictionary imitation (density kernels of libesr_b200.so), periodicity, scribbles and the diverse-solution objectives.  What
is left (VGG / adversarial / desired_SVD / automatic histogram temperature, codes/Z_optimization.py:472-473, :506-508,
:423-425, :479-500) is handed to the reference's own ``Z_optimizer`` class when that module is importable; its
per-iteration ``netG(model_input)`` and ``backward()`` still run on this package's kernels through the generator's
autograd node.  All other names of the reference module are re-exported unchanged."""
import importlib.util as _ilu
import os as _os

from esr_b200 import z_optimization as _b200
from esr_b200.z_optimization import Optimizable_Z, ArcTanH, TV_Loss  # noqa: F401

_spec = _ilu.spec_from_file_location("esr_b200_compat_shadow", _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "_shadow.py"))
_shadow = _ilu.module_from_spec(_spec)
_spec.loader.exec_module(_shadow)
try:
    _reference = _shadow.reexport(__name__, globals())
except Exception as _e:          # e.g. skimage / sklearn missing on this box: only the built objectives are available
    _reference, _reference_error = None, _e

B200_Z_optimizer = _b200.Z_optimizer
Optimizable_Z, ArcTanH, TV_Loss = _b200.Optimizable_Z, _b200.ArcTanH, _b200.TV_Loss


def Z_optimizer(objective, *args, **kwargs):
    """Same call as the reference's class (codes/Z_optimization.py:326-330); returns the optimiser object."""
    from esr_b200 import z_objectives as _zobj
    auto = kwargs.get('auto_set_hist_temperature', args[13] if len(args) > 13 else False)
    if objective in _b200._BUILT or _zobj.unsupported_reason(objective, auto) is None:
        return B200_Z_optimizer(objective, *args, **kwargs)
    if _reference is None:
        raise NotImplementedError("Z objective %r needs the reference's Z_optimization module, which is not importable "
                                  "here (%s)" % (objective, globals().get("_reference_error", "reference tree not on sys.path")))
    return _reference.Z_optimizer(objective, *args, **kwargs)
