"""Synthetic inputs for the reference's own driver scripts: a JSON test file of frames with ground
truth, the dataset's `tm_<a>_<b>.pickle` the scripts derive from its name (test/metrics_from_model.py:109-111), and the
three checkpoint files they load (`pose_estimator.pytorch`, `skeleton_matching.prms`, `skeleton_matching.tch`,
:92-99; formats: skeleton_matching/train_skeleton_matching.py:230-246, pose_estimator/train_pose_estimator.py:269-277).

Deterministic (seeded frames, seeded reference-constructor weights + the stored last-layer calibration of the goldens),
so the build container and the GPU box generate the same files. Nothing here reads the reference tree.
"""
import importlib
import json
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
REPO = os.path.dirname(TESTS)
for p in (REPO, TESTS, os.path.join(HERE, 'stubs')):
    if p not in sys.path:
        sys.path.insert(0, p)

# (seed, persons, synth kwargs): full frames, ragged detections, a frame with one camera only (no graph: skipped by the
# drivers), frames whose first camera is not the first of the rig
FRAME_SPECS = [(300, 4, {}), (301, 3, {}), (302, 4, dict(drop_joint_p=0.15)), (303, 2, {}), (304, 4, dict(pixel_noise=1.0)),
               (305, 3, dict(drop_joint_p=0.3, rand_conf=True)), (306, 4, {}), (307, 3, dict(camera_order=[3, 0, 4, 1])),
               (308, 2, dict(camera_order=[2])), (309, 4, dict(drop_view_p=0.2)), (310, 5, {}), (311, 3, dict(pixel_noise=2.0)),
               (312, 4, {}), (313, 1, {}), (314, 4, dict(drop_joint_p=0.1, drop_view_p=0.1)), (315, 3, {})]
# the viewers (test/show_results_from_model.py:298) stop with torch.cat([]) on a frame whose proposals are empty - with the
# golden weights that is frame 5 - so the viewer file leaves it out; the metrics drivers keep it (their `if batched_input` guard)
VIEWER_SKIP = (5,)


def make(workdir, config='panoptic'):
    import torch
    import helpers
    from pytransform3d.transform_manager import TransformManager
    synth = importlib.import_module('3d_multi_pose_estimator_b200.synth')
    cfg, npz, meta = helpers.load_golden(config)
    os.makedirs(workdir, exist_ok=True)
    models = os.path.join(workdir, 'models')
    os.makedirs(models, exist_ok=True)
    frames = [synth.make_frame(cfg, seed, P, with_gt=True, **kw) for seed, P, kw in FRAME_SPECS]
    test_file = os.path.join(workdir, 'synth_%s_test.json' % config)
    json.dump(frames, open(test_file, 'w'))
    viewer_file = os.path.join(workdir, 'synth_%s_viewer.json' % config)
    json.dump([f for i, f in enumerate(frames) if i not in VIEWER_SKIP], open(viewer_file, 'w'))
    # single-person recordings for test/sm_metrics_without_gt.py (each file holds one individual, :24)
    singles = []
    for k in range(3):
        one = [synth.make_frame(cfg, 400 + 40 * k + t, 1, drop_joint_p=0.1 * (k % 2)) for t in range(12)]
        path = os.path.join(workdir, 'single_%d.json' % k)
        json.dump(one, open(path, 'w'))
        singles.append(path)
    tm = TransformManager()
    for i, n in enumerate(cfg.camera_names):
        tm.add_transform('root', n, cfg.T_root2cam[i])
    tm_file = os.path.join(workdir, 'tm_synth_%s.pickle' % config)
    pickle.dump(tm, open(tm_file, 'wb'))
    gat_state, mlp_state = helpers.golden_weights(config)
    torch.save({'model_state_dict': mlp_state}, os.path.join(models, 'pose_estimator.pytorch'))
    torch.save(gat_state, os.path.join(models, 'skeleton_matching.tch'))
    params = {'loss': 0.0, 'net': 'gat', 'gnn_layers': 5, 'num_feats': cfg.n_features_sm, 'num_hidden': [40, 40, 40, 30],
              'graph_type': '1', 'n_classes': 1, 'heads': [10, 10, 8, 5], 'nonlinearity': torch.nn.LeakyReLU(),
              'final_activation': torch.nn.Sigmoid(), 'in_drop': 0., 'attn_drop': 0., 'alpha': 0.15, 'residual': False}
    pickle.dump(params, open(os.path.join(models, 'skeleton_matching.prms'), 'wb'))
    return dict(test_file=test_file, viewer_file=viewer_file, single_files=singles, frames=frames, tm_dir=workdir, tm_file=tm_file, models=models, n_frames=len(frames))


def make_training_files(workdir, config='panoptic'):
    """Single-person recordings for skeleton_matching/train_skeleton_matching.py (each file one individual, a few frames with
    some views missing): two for --trainset, one each for --devset / --testset. Small on purpose: the script always runs its
    100 epochs (:39), here of one batch each."""
    import helpers
    synth = importlib.import_module('3d_multi_pose_estimator_b200.synth')
    cfg, _, _ = helpers.load_golden(config)
    os.makedirs(workdir, exist_ok=True)
    paths = []
    for k in range(4):
        one = [synth.make_frame(cfg, 600 + 40 * k + t, 1, drop_joint_p=0.1 * (k % 2), drop_view_p=0.25) for t in range(3)]
        path = os.path.join(workdir, 'train_single_%d.json' % k)
        json.dump(one, open(path, 'w'))
        paths.append(path)
    return dict(trainset=paths[:2], devset=paths[2:3], testset=paths[3:4])


if __name__ == '__main__':
    out = make(sys.argv[1] if len(sys.argv) > 1 else '/tmp/b200pose_driver_fixtures')
    out.pop('frames')
    print(out)
