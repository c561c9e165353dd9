"""GPU: the training-side graph path through the drop-in modules, against goldens produced by the
unmodified reference (tests/golden/make_golden_training.py): MergedMultipleHumansDataset over single-person files in
modes 'test_generated' and 'train' (process_training, graph_generator.py:672-810), the collate of
test/sm_metrics_without_gt.py:46-64 with `dgl.batch`, GAT2 forward on single and batched graphs and
get_person_proposal_from_network_output on every graph. Forward only."""
import importlib
import json
import os
import random

import numpy as np
import pytest
import torch

import dropin_env
import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _grad_disabled_like_the_drivers():
    """Every driver these tests restate starts with torch.set_grad_enabled(False) (test/metrics_from_model.py:54,
    test/sm_metrics_without_gt.py:43, ...); with grad enabled the drop-in GAT2 - like the reference's - returns a tensor that
    requires grad (the training path)."""
    prev = torch.is_grad_enabled()
    torch.set_grad_enabled(False)
    yield
    torch.set_grad_enabled(prev)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def collate(batch, dgl, device):
    """test/sm_metrics_without_gt.py:46-64, verbatim in behaviour."""
    graphs = [batch[0][0]]
    batched_labels = batch[0][1]
    batched_indices = batch[0][2]
    total_nodes = batch[0][0].number_of_nodes()
    for graph, labels, indices, nodes_camera in batch[1:]:
        graphs.append(graph)
        batched_labels = torch.cat([batched_labels, labels], dim=0)
        batched_indices = torch.cat([batched_indices, indices + total_nodes], dim=0)
        total_nodes += graph.number_of_nodes()
    return dgl.batch(graphs).to(torch.device(device)), batched_labels, batched_indices


@pytest.mark.parametrize('mode', ['test_generated', 'train'])
def test_training_dataset_batch_and_forward_against_reference(mode, tmp_path):
    cfg, npz, meta = helpers.load_golden('panoptic')
    gz = np.load(os.path.join(GOLDEN, 'golden_training_panoptic.npz'))
    gm = json.load(open(os.path.join(GOLDEN, 'golden_training_panoptic.json')))
    mods = dropin_env.activate(cfg)
    dgl = importlib.import_module('dgl')
    assert dgl.__file__.startswith(dropin_env.SHADOW)
    gg, smu = mods['graph_generator'], mods['skeleton_matching_utils']
    gat_state, _ = helpers.golden_weights('panoptic')
    device = torch.device('cuda')
    model = mods['gat2'].GAT2(None, 5, cfg.n_features_sm, 1, [40, 40, 40, 30], [10, 10, 8, 5], torch.nn.LeakyReLU(),
                              torch.nn.Sigmoid(), 0., 0., 0.15, False, bias=True)
    model.load_state_dict(gat_state)
    model = model.to(device)
    paths = []
    for i, frames in enumerate(gm['files']):
        p = tmp_path / ('single_%d.json' % i)
        p.write_text(json.dumps(frames))
        paths.append(str(p))
    random.seed(gm['seed'])
    ds = gg.MergedMultipleHumansDataset(paths, gm['probabilities'], limit=gm['limit'], mode=mode, alt='3', raw_dir='.',
                                        verbose=False, debug=True)
    recs = gm['modes'][mode]['graphs']
    assert len(ds) == len(recs)
    names = cfg.used_sm_names
    n_props = 0
    for i, rec in enumerate(recs):
        g, labels, indices, nodes_camera = ds[i]
        pre = '%s/%d/' % (mode, i)
        src, dst = [x.cpu().numpy() for x in g.edges()]
        assert np.array_equal(src, gz[pre + 'src']) and np.array_equal(dst, gz[pre + 'dst'])
        assert g.number_of_nodes() == rec['n_nodes']
        assert labels.dtype == torch.float64 and np.array_equal(labels.numpy(), gz[pre + 'labels'])
        assert indices.dtype == torch.int64 and indices.reshape(-1).tolist() == list(range(rec['n_heads'], rec['n_nodes']))
        assert nodes_camera == ['' if c < 0 else names[c] for c in gz[pre + 'nodes_camera']]
        feats = g.ndata['h']
        assert np.array_equal(feats.cpu().numpy().astype(np.float64).sum(axis=1), gz[pre + 'feat_sum'])
        if i == 0:
            assert np.array_equal(feats.cpu().numpy(), gz[pre + 'feats'])
        H, M = rec['n_heads'], rec['n_nodes'] - rec['n_heads']
        assert g.edata['rel_type'].tolist() == [0] * H + [1, 1, 1, 1, 2] * M
        # forward + proposals, as the drivers call them (sm_metrics_without_gt.py:117-135)
        model.g = g
        for layer in model.layers:
            layer.g = g
        outputs = torch.squeeze(model(feats.float(), g))
        ref = gz[pre + 'scores']
        idx = np.arange(H, H + M)
        got = outputs.cpu().numpy()
        assert (np.abs(got[idx] - ref[idx]) / np.abs(ref[idx])).max() <= 1e-4, (mode, i)
        props = smu.get_person_proposal_from_network_output(torch.from_numpy(ref), g, torch.squeeze(indices), nodes_camera, None, 0.5)
        arr = np.array([[-1 if p[c] is None else p[c] for c in names] for p in props], dtype=np.int32).reshape(-1, len(names))
        assert np.array_equal(arr, gz[pre + 'proposals']), (mode, i)
        n_props += len(arr)
    assert n_props > 0
    # ---- the loop body of test/sm_metrics_without_gt.py:107-135: DataLoader batch of ONE graph through collate, forward,
    # proposals on the (one-member) batched graph ----
    for i in range(min(3, len(recs))):
        sg, _, sidx = collate([ds[i]], dgl, device)
        pre = '%s/%d/' % (mode, i)
        model.g = sg
        for layer in model.layers:
            layer.g = sg
        out1 = torch.squeeze(model(sg.ndata['h'].to(device).float(), sg))
        assert sg.nodes().tolist() == list(range(recs[i]['n_nodes']))
        props = smu.get_person_proposal_from_network_output(torch.from_numpy(gz[pre + 'scores']), sg, torch.squeeze(sidx), ds[i][3], None, 0.5)
        arr = np.array([[-1 if p[c] is None else p[c] for c in names] for p in props], dtype=np.int32).reshape(-1, len(names))
        assert np.array_equal(arr, gz[pre + 'proposals']), (mode, i)
        idx = np.arange(recs[i]['n_heads'], recs[i]['n_nodes'])
        ref = gz[pre + 'scores']
        assert (np.abs(out1.cpu().numpy()[idx] - ref[idx]) / np.abs(ref[idx])).max() <= 1e-4
    # ---- dgl.batch through the drivers' collate ----
    nb = gm['modes'][mode]['batch']
    bg, blabels, bindices = collate([ds[i] for i in range(nb)], dgl, device)
    src, dst = [x.cpu().numpy() for x in bg.edges()]
    assert np.array_equal(src, gz['%s/batch/src' % mode]) and np.array_equal(dst, gz['%s/batch/dst' % mode])
    assert bg.batch_size == nb and bg.number_of_nodes() == sum(r['n_nodes'] for r in recs[:nb])
    assert bg.batch_num_nodes().tolist() == [r['n_nodes'] for r in recs[:nb]]
    rel = []
    for r in recs[:nb]:
        rel += [0] * r['n_heads'] + [1, 1, 1, 1, 2] * (r['n_nodes'] - r['n_heads'])
    assert bg.edata['rel_type'].tolist() == rel
    model.g = bg
    for layer in model.layers:
        layer.g = bg
    bfeats = bg.ndata['h']
    bout = torch.squeeze(model(bfeats.float(), bg)).cpu().numpy()
    ref = gz['%s/batch/scores' % mode]
    bi = bindices.reshape(-1).numpy()
    assert (np.abs(bout[bi] - ref[bi]) / np.abs(ref[bi])).max() <= 1e-4
    with pytest.raises(NotImplementedError):
        smu.get_person_proposal_from_network_output(torch.from_numpy(ref), bg, bindices, [], None, 0.5)
    with pytest.raises(AttributeError):
        dgl.graph


def test_many_person_training_graph_and_mixed_batch_vs_oracle():
    """A 12-sample tuple (about 60 heads, ordered pairs: ~2900 edge-nodes, in-degree ~100 - the large-frame aggregation
    kernel and the 256-thread clustering plan on an explicit-list graph) against the oracle restatement; then that graph,
    a test-mode graph and a small training graph in one dgl.batch: member scores unchanged by the batching."""
    from oracle import pose_oracle as O
    cfg, npz, meta = helpers.load_golden('panoptic')
    mods = dropin_env.activate(cfg)
    dgl = importlib.import_module('dgl')
    gg, smu, rt = mods['graph_generator'], mods['skeleton_matching_utils'], mods['rt']
    tg = rt.training_graphs
    ctx = rt.context()
    gat_state, _ = helpers.golden_weights('panoptic')
    model = mods['gat2'].GAT2(None, 5, cfg.n_features_sm, 1, [40, 40, 40, 30], [10, 10, 8, 5], torch.nn.LeakyReLU(),
                              torch.nn.Sigmoid(), 0., 0., 0.15, False, bias=True)
    model.load_state_dict(gat_state)
    model = model.to('cuda')
    synth = helpers.synth
    samples = [synth.make_frame(cfg, 31000 + i, 1 + (i % 3 == 0), drop_joint_p=0.2) for i in range(12)]
    pb, pairs, labels = tg.training_graph_inputs(samples, cfg)
    assert pb.n_heads > 48 and pb.max_heads > pb.n_heads               # in-degree above the head count: ordered pairs
    db = rt.pipeline.HostBatch(pb).to_device(ctx.device)
    arrays = ctx.build_graph_pairs(db, pairs, with_coo=True)
    big = gg.B200Graph(db, arrays, ctx.node_features_f32(db))
    og = O.build_training_graph(samples, O.CameraTables(cfg))
    src, dst = [x.cpu().numpy() for x in big.edges()]
    assert np.array_equal(src, og['src']) and np.array_equal(dst, og['dst'])
    assert np.array_equal(big.ndata['h'].cpu().numpy(), og['feats']) and np.array_equal(labels, og['labels'])
    row_ptr, col = arrays.row_ptr.cpu().numpy(), arrays.col.cpu().numpy()
    order = np.argsort(og['dst'], kind='stable')                        # CSR by destination, in-edges in ascending edge id
    assert np.array_equal(col[: len(order)], og['src'][order])
    assert np.array_equal(np.diff(row_ptr), np.bincount(og['dst'], minlength=og['n_nodes']))

    def forward(g):
        model.set_g(g)
        return torch.squeeze(model(g.ndata['h'].float(), g)).cpu().numpy()
    s_big = forward(big)
    ref = O.gat_forward(helpers.np_state(gat_state), og['feats'], og['src'], og['dst'])
    idx = og['indices']
    assert (np.abs(s_big[idx] - ref[idx]) / np.abs(ref[idx])).max() <= 1e-4
    rng = np.random.default_rng(3)
    H = og['n_heads']
    for sc in (ref, rng.random(len(ref)).astype(np.float32), (0.5 + 0.5 * rng.random(len(ref))).astype(np.float32),
               (np.round(rng.random(len(ref)) * 4) / 4).astype(np.float32)):
        props = smu.get_person_proposal_from_network_output(torch.from_numpy(sc), big, None, None, None, 0.5)
        arr = np.array([[-1 if p[c] is None else p[c] for c in cfg.used_sm_names] for p in props], dtype=np.int32).reshape(-1, cfg.V_sm)
        assert np.array_equal(arr, O.cluster(sc, og['pairs'], og['nodes_camera'][:H], cfg.V_sm, H))
    # ---- mixed batch: explicit-list graphs and a closed-form test-mode graph ----
    frame = meta['frames'][helpers.graph_cases('panoptic')[0]]
    pi = {c: [frame[c][0], 0.0] for c in frame if json.loads(frame[c][0])}
    test_g = gg.MergedMultipleHumansDataset(pi, mode='test', limit=10000, debug=True, alt='3', verbose=False).graphs[0]
    small = tg.training_graph_inputs(samples[:2], cfg)
    sdb = rt.pipeline.HostBatch(small[0]).to_device(ctx.device)
    small_g = gg.B200Graph(sdb, ctx.build_graph_pairs(sdb, small[1]), ctx.node_features_f32(sdb))
    members = [small_g, test_g, big]
    singles = [forward(g) for g in members]
    bg = dgl.batch(members)
    sb = forward(bg)
    assert bg.number_of_nodes() == sum(g.number_of_nodes() for g in members)
    off = 0
    bsrc, bdst = [x.cpu().numpy() for x in bg.edges()]
    e_off = 0
    for g, s in zip(members, singles):
        n, e = g.number_of_nodes(), g.number_of_edges()
        gs, gd = [x.cpu().numpy() for x in g.edges()]
        assert np.array_equal(bsrc[e_off:e_off + e], gs + off) and np.array_equal(bdst[e_off:e_off + e], gd + off)
        assert (np.abs(sb[off:off + n] - s) / np.abs(s)).max() <= 2e-5
        off += n
        e_off += e
