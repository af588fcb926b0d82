"""Importable alias of the ``differentiable-ilqr_b200`` package (whose directory
name contains a hyphen): ``import dilqr_b200 as dilqr``."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("differentiable-ilqr_b200")
sys.modules[__name__] = _pkg
