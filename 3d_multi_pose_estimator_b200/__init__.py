"""B200-native per-frame inference hot path of gnns4hri/3D_multi_pose_estimator.

Package name starts with a digit, so import it with
`importlib.import_module("3d_multi_pose_estimator_b200")` (the top-level `b200pose.py`
alias does exactly that).
"""
from .config import CameraConfig, ring_config, N_JOINTS  # noqa: F401

__version__ = "0.1.0"
