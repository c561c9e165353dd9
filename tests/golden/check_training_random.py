"""One-off pinning run (needs /root/reference; not part of the pytest suites): the oracle's restatement of process_training
against the UNMODIFIED reference on 120 random tuples of single-person samples - spurious skeletons, empty cameras,
skeletons without joints, single-camera samples - the same generator tests/test_training_graphs_cpu.py uses for the
product's host logic. Every file holds one sample, and the seed is chosen so that the reference's sampler takes all files,
i.e. the dataset yields exactly the tuple's graph.

    python tests/golden/check_training_random.py      # prints the number of graphs compared; exits non-zero on a mismatch
"""
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(REPO, 'tests'))
import make_golden as MG  # noqa: E402


def random_tuple(cfg, synth, rng, case):
    samples = []
    for s in range(int(rng.integers(1, 6))):
        fr = synth.make_frame(cfg, 40000 + 10 * case + s, 1, drop_joint_p=float(rng.uniform(0, 0.6)),
                              drop_view_p=float(rng.uniform(0, 0.7)), rand_conf=True, keep_empty=bool(rng.integers(0, 2)))
        for cam in list(fr):
            r = rng.random()
            if r < 0.2:
                extra = json.loads(synth.make_frame(cfg, 50000 + 10 * case + s, 1, drop_joint_p=0.6)[cam][0])
                sk = json.loads(fr[cam][0])
                fr[cam][0] = json.dumps(extra + sk if rng.random() < 0.5 else sk + extra)
            elif r < 0.3:
                fr[cam][0] = '[]'
            elif r < 0.4:
                del fr[cam]
        samples.append(fr)
    return samples


def main():
    import contextlib
    import io
    import importlib
    import torch
    torch.set_grad_enabled(False)
    parameters = MG._activate('panoptic')
    import graph_generator
    b200 = importlib.import_module('3d_multi_pose_estimator_b200')
    synth = importlib.import_module('3d_multi_pose_estimator_b200.synth')
    from oracle import pose_oracle as O
    cfg = b200.CameraConfig.from_parameters(parameters, name='panoptic')
    tabs = O.CameraTables(cfg)
    names = parameters.used_cameras_skeleton_matching
    rng = np.random.default_rng(11)
    compared = 0
    for case in range(120):
        samples = random_tuple(cfg, synth, rng, case)
        n = len(samples)
        paths = []
        for i, fr in enumerate(samples):
            p = '/tmp/b200pose_rand_%d.json' % i
            json.dump([fr], open(p, 'w'))
            paths.append(p)
        probs = [0.1 + 0.1 * i for i in range(n)]
        seed = next(s for s in range(1000) if random.Random(s).randint(1, n) == n)
        random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            ds = graph_generator.MergedMultipleHumansDataset(paths, probs, limit=1, mode='test_generated', alt='3', raw_dir='.',
                                                             verbose=False, debug=True)
        random.seed(seed)
        inputs, indices = O.load_training_inputs([[json.loads(json.dumps(fr))] for fr in samples], 'test_generated',
                                                 parameters.used_cameras, random)
        og = [O.build_training_graph(mp, tabs) for mp in O.training_samples(inputs, indices, probs, 1, random)]
        og = [g for g in og if g is not None]
        assert len(og) == len(ds.graphs), (case, len(og), len(ds.graphs))
        for g, o in zip(ds.graphs, og):
            src, dst = [t.numpy().astype(np.int32) for t in g.edges()]
            assert np.array_equal(src, o['src']) and np.array_equal(dst, o['dst']), case
            assert np.array_equal(g.ndata['h'].numpy(), o['feats']), case
            assert np.array_equal(ds.labels[0].numpy(), o['labels']), case
            cams = np.array([names.index(c) if c else -1 for c in ds.data['nodes_camera'][0]], dtype=np.int32)
            assert np.array_equal(cams, o['nodes_camera']), case
            compared += 1
    print('process_training: oracle == reference on %d random tuples (%d without a graph)' % (compared, 120 - compared))


if __name__ == '__main__':
    main()
