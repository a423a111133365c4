"""`pantea` import surface of the B200-native implementation: `from pantea.atoms import Structure`,
`from pantea.potentials import NeuralNetworkPotential`, `from pantea.simulation import MDSimulator, simulate`, ...
resolve to the same-named modules of `pantea_b200` (reference `pantea/__init__.py` and the sub-package `__init__`s).
An alias, not a copy: every `pantea.x.y` entry of `sys.modules` IS the `pantea_b200.x.y` module object."""
import importlib
import pkgutil
import sys

import pantea_b200 as _impl

__version__ = getattr(_impl, "__version__", "0.11.0+b200")
_SKIP = ("pantea_b200.csrc", "pantea_b200.variants", "pantea_b200.jax_ffi")

for _info in pkgutil.walk_packages(_impl.__path__, prefix="pantea_b200."):
    if _info.name.startswith(_SKIP):
        continue
    try:
        _mod = importlib.import_module(_info.name)
    except Exception:  # a leaf with an optional dependency (ase, jax)
        continue
    _name = "pantea" + _info.name[len("pantea_b200"):]
    sys.modules[_name] = _mod
    _parent, _, _leaf = _name.rpartition(".")
    if _parent == "pantea":
        globals()[_leaf] = _mod
