"""Importable alias of the package directory `naqs-for-quantum-chemistry_b200/` (a dash cannot appear in an
`import` statement).  `import naqs_b200` gives the very same module object."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)
_real = importlib.import_module("naqs-for-quantum-chemistry_b200")
for _name, _mod in list(sys.modules.items()):
    if _name.startswith("naqs-for-quantum-chemistry_b200."):
        sys.modules["naqs_b200." + _name.split(".", 1)[1]] = _mod
sys.modules[__name__] = _real
