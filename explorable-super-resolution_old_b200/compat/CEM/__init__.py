"""`CEM` package shim: CEM.CEMnet comes from this directory, everything else (imresize_CEM, ...) from the reference."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
