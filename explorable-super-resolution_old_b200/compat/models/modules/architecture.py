"""`models.modules.architecture`: every class of the reference's module (discriminators, VGG extractor, SRResNet,
codes/models/modules/architecture.py) with ``RRDBNet`` (:102-175) replaced by the B200 generator."""
from .._shadow_loader import reexport as _reexport

_reference = _reexport(__name__, globals())
from esr_b200.rrdbnet import RRDBNet  # noqa: E402,F401
from esr_b200.discriminator import Discriminator_VGG_128_  # noqa: E402,F401
