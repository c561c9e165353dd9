"""Shared fixtures/helpers for the parity tests (golden loading, seeded weights)."""
import functools
import importlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, 'golden')
pkg = importlib.import_module('3d_multi_pose_estimator_b200')
weights_mod = importlib.import_module('3d_multi_pose_estimator_b200.weights')
synth = importlib.import_module('3d_multi_pose_estimator_b200.synth')

# panoptic / arp3 / ring10: BASELINE configs 1-2 / 3 / 5; arp6: the reference's shipped ARPLAB rig (narrow-baseline stereo
# pair included); arp_robot2: its robot-cameras-only variant (parameters.py:110-112: used cameras a subset of the rig);
# pansub: Panoptic with a permuted 4-of-5 skeleton-matching set and a 3-camera pose-estimator set
CONFIGS = ['panoptic', 'arp3', 'ring10', 'arp6', 'arp_robot2', 'pansub']


@functools.lru_cache(maxsize=None)
def load_golden(config):
    npz = np.load(os.path.join(GOLDEN, 'golden_%s.npz' % config))
    meta = json.load(open(os.path.join(GOLDEN, 'golden_%s.json' % config)))
    cfg = pkg.CameraConfig.from_npz(os.path.join(GOLDEN, 'cameras_%s.npz' % config))
    return cfg, npz, meta


@functools.lru_cache(maxsize=None)
def golden_weights(config):
    """(gat_state, mlp_state) as torch tensors: seeded reference init + the stored calibration of the
    last GAT layer (tests/golden/make_golden.py:calibrate)."""
    import torch
    cfg, npz, meta = load_golden(config)
    gat = weights_mod.make_gat_state(cfg.n_features_sm, meta['gat_seed'])
    gat['layers.4.fc2.weight'] = torch.from_numpy(npz['gat_last_fc2_weight'].copy())
    gat['layers.4.fc2.bias'] = torch.from_numpy(npz['gat_last_fc2_bias'].copy())
    mlp = weights_mod.make_mlp_state(meta.get('mlp_in_dim', cfg.n_cameras * 18 * 14), 54, meta['mlp_seed'])
    return gat, mlp


def np_state(state):
    return {k: v.numpy() for k, v in state.items()}


def graph_cases(config):
    cfg, npz, meta = load_golden(config)
    return [t for t in meta['cases'] if t not in meta['no_graph']]


def person_dicts(frame_graph, proposals, cfg):
    """{camera_name: skeleton} per person, in used_cameras order (metrics_from_model.py:248-252)."""
    out = []
    sm = cfg.used_sm_names
    for person in proposals:
        p = {}
        for cam in cfg.used_pe_names:
            if cam in sm and person[sm.index(cam)] >= 0:
                p[cam] = frame_graph['heads_json'][int(person[sm.index(cam)])]
        out.append(p)
    return out
