"""CPU: the training-side restatement (oracle) and the product's host-side sample / edge-node logic
(3d_multi_pose_estimator_b200/training_graphs.py) against the goldens the unmodified reference produced
(tests/golden/make_golden_training.py). No GPU: the device part is covered by tests/test_training_graphs_gpu.py."""
import importlib
import json
import os
import random

import numpy as np
import pytest

import helpers
from oracle import pose_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
tg = importlib.import_module('3d_multi_pose_estimator_b200.training_graphs')


def _golden():
    return (np.load(os.path.join(GOLDEN, 'golden_training_panoptic.npz')),
            json.load(open(os.path.join(GOLDEN, 'golden_training_panoptic.json'))))


@pytest.mark.parametrize('mode', ['test_generated', 'train'])
def test_oracle_training_graphs_match_reference(mode):
    cfg, _, _ = helpers.load_golden('panoptic')
    gz, gm = _golden()
    tabs = O.CameraTables(cfg)
    random.seed(gm['seed'])
    inputs, indices = O.load_training_inputs(json.loads(json.dumps(gm['files'])), mode, cfg.used_pe_names, random)
    graphs = [g for g in (O.build_training_graph(mp, tabs) for mp in
                          O.training_samples(inputs, indices, gm['probabilities'], gm['limit'], random)) if g is not None]
    recs = gm['modes'][mode]['graphs']
    assert len(graphs) == len(recs)
    for i, g in enumerate(graphs):
        pre = '%s/%d/' % (mode, i)
        assert np.array_equal(g['src'], gz[pre + 'src']) and np.array_equal(g['dst'], gz[pre + 'dst'])
        assert np.array_equal(g['labels'], gz[pre + 'labels'])
        assert np.array_equal(g['nodes_camera'], gz[pre + 'nodes_camera'])
        assert np.array_equal(g['feats'].astype(np.float64).sum(axis=1), gz[pre + 'feat_sum'])
        props = O.cluster(gz[pre + 'scores'], g['pairs'], g['nodes_camera'][:g['n_heads']], cfg.V_sm, g['n_heads'])
        assert np.array_equal(props, gz[pre + 'proposals'])
    nb = gm['modes'][mode]['batch']
    b = O.batch_graphs(graphs[:nb])
    assert np.array_equal(b['src'], gz['%s/batch/src' % mode]) and np.array_equal(b['dst'], gz['%s/batch/dst' % mode])
    # GAT forward of the restatement on the first graph and on the batch (block-diagonal: same scores per member)
    gat_w = helpers.np_state(helpers.golden_weights('panoptic')[0])
    s0 = O.gat_forward(gat_w, graphs[0]['feats'], graphs[0]['src'], graphs[0]['dst'])
    idx = graphs[0]['indices']
    assert (np.abs(s0[idx] - gz['%s/0/scores' % mode][idx]) / np.abs(gz['%s/0/scores' % mode][idx])).max() <= 1e-5


@pytest.mark.parametrize('mode', ['test_generated', 'train'])
def test_host_sample_and_edge_node_logic_matches_reference(mode):
    cfg, _, _ = helpers.load_golden('panoptic')
    gz, gm = _golden()
    random.seed(gm['seed'])
    inputs, indices = tg.load_inputs(json.loads(json.dumps(gm['files'])), mode, cfg.used_pe_names, random)
    i = 0
    for mp in tg.sample_sets(inputs, indices, gm['probabilities'], gm['limit'], random):
        built = tg.training_graph_inputs(mp, cfg)
        if built is None:
            continue
        pb, pairs, labels = built
        H = pb.n_heads
        src, dst = list(range(H)), list(range(H))
        for k, (a, b) in enumerate(pairs.tolist()):
            e = H + k
            src += [a, e, b, e, e]
            dst += [e, a, e, b, e]
        pre = '%s/%d/' % (mode, i)
        assert src == gz[pre + 'src'].tolist() and dst == gz[pre + 'dst'].tolist()
        assert np.array_equal(labels, gz[pre + 'labels'])
        assert pb.n_frames == 1 and pb.node_off.tolist() == [0, H + len(pairs)]
        assert pb.max_heads >= 1 + np.bincount(pairs.ravel()).max()
        i += 1
    assert i == len(gm['modes'][mode]['graphs'])


def test_augment_views_subsets():
    """data_augmentation.py:50-89: the full view set first, then its proper subsets with >= min_views cameras."""
    used = ['a', 'b', 'c']
    sample = {'x': ['[{"0": [0, 1, 1, 1, 1]}]'], 'a': ['[{"0": [0, 1, 1, 1, 1]}]'], 'b': ['[]'], 'c': ['[{"1": [1, 2, 2, 1, 1]}]']}
    out = tg.augment_views([sample], used, 1)
    assert [list(o) for o in out] == [['a', 'c'], ['c'], ['a']]
    assert [list(o) for o in tg.augment_views([sample], used, 2)] == [['a', 'c']]
    assert tg.augment_views([{'b': ['[]']}], used, 1) == []
    assert [list(o) for o in O.augment_views([sample], used, 1)] == [['a', 'c'], ['c'], ['a']]


def test_edge_node_list_matches_oracle_on_random_tuples():
    """The product's host-side edge-node / label logic against the oracle restatement of process_training on 120 random
    tuples of 1..5 single-person samples with spurious skeletons, empty cameras, skeletons without joints and
    single-camera samples (graph_generator.py:699-800)."""
    import json as _json
    cfg, _, _ = helpers.load_golden('panoptic')
    tabs = O.CameraTables(cfg)
    rng = np.random.default_rng(11)
    n_graphs = 0
    for case in range(120):
        samples = []
        for s in range(int(rng.integers(1, 6))):
            fr = helpers.synth.make_frame(cfg, 40000 + 10 * case + s, 1, drop_joint_p=float(rng.uniform(0, 0.6)),
                                          drop_view_p=float(rng.uniform(0, 0.7)), rand_conf=True, keep_empty=bool(rng.integers(0, 2)))
            for cam in list(fr):
                r = rng.random()
                if r < 0.2:                                     # spurious partial skeleton(s) in this camera
                    extra = _json.loads(helpers.synth.make_frame(cfg, 50000 + 10 * case + s, 1, drop_joint_p=0.6)[cam][0])
                    sk = _json.loads(fr[cam][0])
                    fr[cam][0] = _json.dumps(extra + sk if rng.random() < 0.5 else sk + extra)
                elif r < 0.3:
                    fr[cam][0] = '[]'                           # camera present but empty
                elif r < 0.4:
                    del fr[cam]
            samples.append(fr)
        og = O.build_training_graph(samples, tabs)
        built = tg.training_graph_inputs(samples, cfg)
        assert (og is None) == (built is None), case
        if og is None:
            continue
        pb, pairs, labels = built
        assert pb.n_heads == og['n_heads']
        assert np.array_equal(pairs, og['pairs']) and np.array_equal(labels, og['labels']), case
        assert np.array_equal(pb.sk_cam, np.array([cfg.used_sm[c] for c in og['nodes_camera'][:og['n_heads']]], dtype=np.int32))
        assert pb.max_heads >= 1 + np.bincount(pairs.ravel()).max() and pb.max_enodes == len(pairs)
        n_graphs += 1
    assert n_graphs > 60
