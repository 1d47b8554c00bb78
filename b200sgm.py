"""Import alias: the package directory is named `i3dr_stereo_camera-ros_b200` (not a Python identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("i3dr_stereo_camera-ros_b200")
sys.modules[__name__] = _pkg
