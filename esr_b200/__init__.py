"""Importable alias for the product package.

The product lives in ``explorable-super-resolution_old_b200/`` (a directory name
Python cannot import directly); this package forwards its search path there, so
``import esr_b200.cem`` loads ``explorable-super-resolution_old_b200/cem.py``.
"""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "explorable-super-resolution_old_b200")
__path__.insert(0, _impl)
PACKAGE_DIR = _impl
