"""B200-native per-frame inference hot path of gnns4hri/3D_multi_pose_estimator.

Package name starts with a digit, so import it with
`importlib.import_module("3d_multi_pose_estimator_b200")` (the top-level `b200pose.py`
alias does exactly that).
"""
import os as _os


def _widen_stream_queues():
    """The streamed host path runs on up to ten CUDA streams (compute lanes and their graph-builder side streams, the copy-in
    stream, two read-back streams, NCCL's own). CUDA maps streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues - 8 by
    default - and streams that share a queue pick up each other's waits: with three lanes on a multi-GPU box (NCCL's streams
    shift the mapping) the read-back stream's wait for batch i landed in front of another lane's kernels and a step took 3.9 ms
    instead of 2.0. The variable is read when the process creates its CUDA context, so it is set here, at import, unless the user
    chose a value or the context already exists (then `lanes_ok` says whether more than two lanes are safe)."""
    have = _os.environ.get('CUDA_DEVICE_MAX_CONNECTIONS')
    if have is not None:
        try:
            return int(have) >= 16
        except ValueError:
            return False
    try:
        import torch
        if torch.cuda.is_initialized():
            return False
    except Exception:
        pass
    _os.environ['CUDA_DEVICE_MAX_CONNECTIONS'] = '32'
    return True


lanes_ok = _widen_stream_queues()
DEFAULT_LANES = 3 if lanes_ok else 2

from .config import CameraConfig, ring_config, N_JOINTS  # noqa: E402,F401

__version__ = "0.1.0"
