"""`import models.networks as networks`: the reference's factories (define_D, define_F, ... codes/models/networks.py)
with ``define_G`` / ``init_weights`` / ``weights_init_kaiming`` (:28-102) replaced by the B200 generator factory and
``define_D`` (:105-127) by the one that can construct the critic (the shipped one passes ``nb=`` to a class that does not
take it; esr_b200/discriminator.py)."""
from ._shadow_loader import reexport as _reexport

try:
    _reference = _reexport(__name__, globals())
except Exception as _e:      # the reference's own module failed to import (missing third-party package): generator only
    _reference, _reference_error = None, _e
from esr_b200.networks import define_G, init_weights, weights_init_kaiming  # noqa: E402,F401
from esr_b200.discriminator import define_D  # noqa: E402,F401

if _reference is None:
    def define_F(*a, **kw):
        raise NotImplementedError("models.networks.define_F: the reference tree is not importable here and the VGG "
                                  "feature extractor is outside this package's hot path")
