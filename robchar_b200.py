"""Importable alias for the package directory ``code-robchar_b200/`` (a hyphen cannot appear in an
``import`` statement).  ``import robchar_b200`` returns that package; its submodules are registered
under both names so they are loaded exactly once."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("code-robchar_b200")
for _name, _mod in list(sys.modules.items()):
    if _name.startswith("code-robchar_b200."):
        sys.modules["robchar_b200." + _name.split(".", 1)[1]] = _mod
sys.modules[__name__] = _pkg
