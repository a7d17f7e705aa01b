"""Import alias: the package directory is named after the reference repo
(`decision-making-and-path-planning_b200`), which is not a valid Python identifier."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("decision-making-and-path-planning_b200")
sys.modules[__name__] = _pkg
