"""Short import alias for the package directory (whose mandated name is not a Python identifier).

``import lewin_b200`` (and ``lewin_b200.ops`` etc.) resolve to the SAME module objects as the package
``research-and-implementation-of-image-dehazing-algorithm-based-on-vision-transformer_b200``.
"""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_LONG = "research-and-implementation-of-image-dehazing-algorithm-based-on-vision-transformer_b200"
_pkg = importlib.import_module(_LONG)
for _sub in ("_lib", "options", "ops", "modules", "patch", "uformer", "fullres", "parallel", "losses", "training", "canvas_bands"):
    importlib.import_module(f"{_LONG}.{_sub}")
for _name, _mod in list(sys.modules.items()):
    if _name.startswith(_LONG + "."):
        sys.modules[__name__ + _name[len(_LONG):]] = _mod
sys.modules[__name__] = _pkg
