"""`models.modules` package shim: architecture comes from this directory, block / loss / spectral_norm / archs_util
from the reference."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
