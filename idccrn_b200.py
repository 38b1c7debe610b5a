"""Import alias: the package directory is ``i-dccrn-vae_b200/`` (not a Python identifier), so
``import idccrn_b200`` resolves to it here.  The package imports every one of its submodules in ``__init__`` and all of
them are aliased below, so there is exactly one module object (one copy of every class and of every module-level
switch) regardless of which name was used to import it."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)

_REAL = "i-dccrn-vae_b200"
_pkg = importlib.import_module(_REAL)
for _k in list(sys.modules):
    if _k.startswith(_REAL + "."):
        sys.modules[__name__ + _k[len(_REAL):]] = sys.modules[_k]
sys.modules[__name__] = _pkg
