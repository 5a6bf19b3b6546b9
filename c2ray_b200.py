"""Import shim: `import c2ray_b200` -> the package directory `c2-ray3dm1d_helium_b200/` (not a valid identifier)."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
_pkg = importlib.import_module("c2-ray3dm1d_helium_b200")
sys.modules[__name__] = _pkg
