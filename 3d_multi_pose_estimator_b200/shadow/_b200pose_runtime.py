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

live_mod = importlib.import_module('3d_multi_pose_estimator_b200.live')

_cfg = None
_ctx = None
_parameters = None
_live = None
_active = {'gat': None, 'mlp': None}        # (module id, weights key, prepared planes) of the models the driver runs
LIVE_ENABLED = os.environ.get('B200POSE_DROPIN_LIVE', '1') != '0'


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
    global _cfg, _ctx, _live
    _cfg = cfg
    _ctx = None
    _live = None
    _active['gat'] = _active['mlp'] = None


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


# ---- live frames (3d_multi_pose_estimator_b200/live.py): the drop-in dataset submits a frame once, the later calls of the
# reference's driver loop are answered from that submission ---------------------------------------------------------------
def note_model(kind, module, key, prepared):
    """Called by GAT2.forward / PoseEstimatorMLP.forward with the weights they just ran: from then on the dataset drop-in
    can submit whole frames with them."""
    _active[kind] = (id(module), key, prepared)


def live():
    """The LiveFrames of the context, bound to the models the driver has been running; None until a GAT2 was seen."""
    global _live
    if not LIVE_ENABLED or _active['gat'] is None:
        return None
    ctx = context()
    if _live is None:
        _live = live_mod.LiveFrames(ctx)
    g, m = _active['gat'], _active['mlp']
    _live.set_models((g[0], g[1], None if m is None else (m[0], m[1])), g[2], None if m is None else m[2])
    return _live


def last_live():
    return _live.last if _live is not None else None


def sync_live():
    """Before eager kernels on the shared pipeline: order the current stream after the submitted frames."""
    if _live is not None:
        _live.sync()
