"""Shared runtime of the drop-in modules in this directory.

The reference's scripts find their modules through `sys.path.append('../skeleton_matching' | '../utils' | '../')`
with cwd = <reference>/test (test/metrics_from_model.py:12,17,23). Put THIS directory earlier on sys.path
(e.g. PYTHONPATH=<repo>/3d_multi_pose_estimator_b200/shadow) and the same imports resolve to the B200 path:
gat2, graph_generator, mlp, pose_estimator_utils, skeleton_matching_utils, pose_estimator_dataset_from_json.
`parameters` (and the pickled TransformManager it points to) stay the reference's own.

Configuration lookup order: an explicit `set_config(CameraConfig)` (tests, embedding applications), else the
reference's `parameters.parameters` namedtuple + pickle exactly as the reference modules read them at import
time (skeleton_matching/graph_generator.py:24,32; utils/pose_estimator_dataset_from_json.py:4,28).
"""
import importlib
import os
import sys

_REPO = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

pkg = importlib.import_module('3d_multi_pose_estimator_b200')
pipeline = importlib.import_module('3d_multi_pose_estimator_b200.pipeline')
pack = importlib.import_module('3d_multi_pose_estimator_b200.pack')
training_graphs = importlib.import_module('3d_multi_pose_estimator_b200.training_graphs')

_cfg = None
_ctx = None
_parameters = None


def parameters():
    """The reference's `parameters.parameters` namedtuple (looked up like the reference modules do)."""
    global _parameters
    if _parameters is None:
        if 'parameters' not in sys.modules and '../' not in sys.path:
            sys.path.append('../')
        _parameters = importlib.import_module('parameters').parameters
    return _parameters


def set_config(cfg):
    """Use an explicit CameraConfig instead of the reference's `parameters` module. Resets the device context."""
    global _cfg, _ctx
    _cfg = cfg
    _ctx = None


def config():
    """CameraConfig built from the reference's `parameters` + pickle, once."""
    global _cfg
    if _cfg is None:
        _cfg = pkg.CameraConfig.from_parameters(parameters())
    return _cfg


def context():
    """The process-wide PosePipeline (camera tables + workspaces on the current CUDA device).
    Raises if there is no CUDA device: the drop-in has no CPU fallback."""
    global _ctx
    if _ctx is None:
        _ctx = pipeline.PosePipeline(config(), None, None)
    return _ctx
