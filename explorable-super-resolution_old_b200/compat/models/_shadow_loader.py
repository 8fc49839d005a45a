"""Gives the shims inside the `models` package access to compat/_shadow.py (which is a top-level module of the
compat directory, not part of any package)."""
import importlib.util
import os

_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_shadow.py")
_spec = importlib.util.spec_from_file_location("esr_b200_compat_shadow", _path)
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
load_shadowed, reexport = _mod.load_shadowed, _mod.reexport
