"""Test helper: makes the drop-in modules of 3d_multi_pose_estimator_b200/shadow importable under their
reference names, with a stand-in `parameters` module built from a golden CameraConfig (the reference tree,
which normally provides `parameters.py`, does not exist on the GPU box)."""
import collections
import importlib
import os
import sys
import types

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHADOW = os.path.join(REPO, '3d_multi_pose_estimator_b200', 'shadow')
MODULES = ['gat2', 'graph_generator', 'mlp', 'pose_estimator_utils', 'skeleton_matching_utils',
           'pose_estimator_dataset_from_json']


def stub_parameters(cfg):
    """A namedtuple with the fields of the reference's parameters.py:10-43 that the hot path reads."""
    f = lambda a: [float(x) for x in a]
    fields = dict(image_width=int(cfg.image_width), image_height=int(cfg.image_height),
                  cameras=list(range(cfg.n_cameras)), camera_names=list(cfg.camera_names),
                  fx=f(cfg.fx), fy=f(cfg.fy), cx=f(cfg.cx), cy=f(cfg.cy), kd0=f(cfg.kd0), kd1=f(cfg.kd1), kd2=f(cfg.kd2),
                  p1=f(cfg.p1), p2=f(cfg.p2), joint_list=list(range(18)), numbers_per_joint=14, numbers_per_joint_for_loss=4,
                  transformations_path=None, used_cameras=cfg.used_pe_names, used_cameras_skeleton_matching=cfg.used_sm_names,
                  used_joints=[0, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17], min_number_of_views=cfg.min_number_of_views,
                  format='COCO', graph_alternative='3',
                  axes_3D={'Y': (cfg.median_axis, 1.), 'Z': (cfg.up_axis, cfg.up_sign), 'X': (0, 1.)})
    T = collections.namedtuple('TrackerParameters', list(fields))
    return T(**fields)


def activate(cfg):
    """Returns {name: module} of the six drop-in modules, configured for `cfg`."""
    if SHADOW not in sys.path:
        sys.path.insert(0, SHADOW)
    mod = types.ModuleType('parameters')
    mod.parameters = stub_parameters(cfg)
    sys.modules['parameters'] = mod
    rt = importlib.import_module('_b200pose_runtime')
    rt._parameters = mod.parameters
    rt.set_config(cfg)
    out = {}
    for name in MODULES:
        if name in sys.modules and getattr(sys.modules[name], '__file__', '').startswith(SHADOW):
            out[name] = importlib.reload(sys.modules[name])
        else:
            sys.modules.pop(name, None)
            out[name] = importlib.import_module(name)
    out['rt'] = rt
    return out
