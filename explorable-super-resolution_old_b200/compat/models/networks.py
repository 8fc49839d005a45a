"""`import models.networks as networks` -> B200 generator factory (reference: codes/models/networks.py)."""
from esr_b200.networks import define_G, init_weights, weights_init_kaiming  # noqa: F401


def define_D(*a, **kw):
    raise NotImplementedError("the discriminator / GAN training step is outside the built hot path (SURVEY.md §8f)")


def define_F(*a, **kw):
    raise NotImplementedError("the VGG feature extractor is outside the built hot path (SURVEY.md §8f)")
