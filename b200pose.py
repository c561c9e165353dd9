"""Import alias: `import b200pose` == importlib.import_module('3d_multi_pose_estimator_b200')."""
import importlib
import sys

_pkg = importlib.import_module("3d_multi_pose_estimator_b200")
sys.modules[__name__] = _pkg
