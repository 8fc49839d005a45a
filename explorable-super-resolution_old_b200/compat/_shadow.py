"""Finds the reference's module that a same-named shim in this directory shadows.

``compat/`` sits AHEAD of the reference's ``codes/`` on ``sys.path``; its packages extend their ``__path__`` over
every later ``sys.path`` entry (``pkgutil.extend_path``), so modules that exist only in the reference
(``models.SRRaGAN_model``, ``models.base_model``, ``models.modules.loss`` / ``block``, ``CEM.imresize_CEM``, ...) keep
resolving to the reference, while the few modules that exist here win.  A shim that wants to keep everything its
reference namesake defines (discriminators, other generators, loss helpers) loads that namesake through
``load_shadowed`` and re-exports its names before overriding the hot-path ones."""
import importlib.util
import os
import sys

COMPAT_DIR = os.path.dirname(os.path.abspath(__file__))


def load_shadowed(module_name):
    """The module `module_name` as found in the first ``sys.path`` entry other than this directory, or None when the
    reference tree is not on the path.  It is registered as ``<module_name>__reference`` (relative imports inside it
    resolve against the same parent package)."""
    alias = module_name + "__reference"
    if alias in sys.modules:
        return sys.modules[alias]
    rel = module_name.replace(".", os.sep) + ".py"
    for entry in sys.path:
        base = os.path.abspath(entry or os.getcwd())
        if base == COMPAT_DIR:
            continue
        path = os.path.join(base, rel)
        if os.path.isfile(path):
            parent = module_name.rpartition(".")[0]
            spec = importlib.util.spec_from_file_location(alias, path)
            mod = importlib.util.module_from_spec(spec)
            mod.__package__ = parent
            sys.modules[alias] = mod
            try:
                spec.loader.exec_module(mod)
            except BaseException:
                del sys.modules[alias]
                raise
            return mod
    return None


def reexport(module_name, namespace):
    """Copies the public names of the shadowed reference module into `namespace`; returns the module (or None)."""
    ref = load_shadowed(module_name)
    if ref is not None:
        for k, v in vars(ref).items():
            if not k.startswith("__"):
                namespace.setdefault(k, v)
    return ref
