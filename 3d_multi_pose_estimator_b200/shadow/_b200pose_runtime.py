"""Shared runtime of the drop-in modules in this directory.

The reference's scripts find their modules through `sys.path.append('../skeleton_matching' | '../utils' | '../')`
with cwd = <reference>/test (test/metrics_from_model.py:12,17,23). Put THIS directory earlier on sys.path
(e.g. PYTHONPATH=<repo>/3d_multi_pose_estimator_b200/shadow) and the same imports resolve to the B200 path:
gat2, graph_generator, mlp, pose_estimator_utils, skeleton_matching_utils, pose_estimator_dataset_from_json.
`parameters` (and the pickled TransformManager it points to) stay the reference's own.
"""
import importlib
import os
import sys

_REPO = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)
sys.path.append('../')                      # same lookup the reference modules do for `parameters`
from parameters import parameters           # noqa: E402

pkg = importlib.import_module('3d_multi_pose_estimator_b200')
pipeline = importlib.import_module('3d_multi_pose_estimator_b200.pipeline')
pack = importlib.import_module('3d_multi_pose_estimator_b200.pack')

_cfg = None
_ctx = None


def config():
    """CameraConfig built from the reference's `parameters` + pickle, once."""
    global _cfg
    if _cfg is None:
        _cfg = pkg.CameraConfig.from_parameters(parameters)
    return _cfg


def context():
    """The process-wide PosePipeline (camera tables + workspaces on the current CUDA device).
    Raises if there is no CUDA device: the drop-in has no CPU fallback."""
    global _ctx
    if _ctx is None:
        _ctx = pipeline.PosePipeline(config(), None, None)
    return _ctx
