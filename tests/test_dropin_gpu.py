"""GPU: the reference's per-frame driver loop, written against the drop-in modules under their reference names
(call sequence of test/metrics_from_model.py:178-300 and test/metrics_from_triangulation.py:234-249), compared
with the goldens the unmodified reference produced for the same frames."""
import json

import numpy as np
import pytest
import torch

import dropin_env
import helpers
from oracle import check as OC
from oracle import pose_oracle as O

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


def run_frame(mods, cfg, model, mlp, frame):
    """One pass of the reference driver's loop body. Returns None when the frame yields no graph."""
    gg, smu, ds = mods['graph_generator'], mods['skeleton_matching_utils'], mods['pose_estimator_dataset_from_json']
    device = torch.device('cuda')
    processed_input = {}
    for cam in frame:                                                    # metrics_from_model.py:182-191
        cam_data = json.loads(frame[cam][0])
        if cam_data:
            processed_input[cam] = [json.dumps(cam_data), frame[cam][1]]
    scenario = gg.MergedMultipleHumansDataset(processed_input, mode='test', limit=10000, debug=True, alt='3', verbose=False)
    if len(scenario.graphs) == 0:
        return None
    subgraph = scenario.graphs[0].to(device)
    indices = scenario.data['edge_nodes_indices'][0].to(device)
    nodes_camera = scenario.data['nodes_camera'][0]
    feats = subgraph.ndata['h'].to(device)
    model.g = subgraph
    for layer in model.layers:
        layer.g = subgraph
    outputs = torch.squeeze(model(feats.float(), subgraph))
    indices = torch.squeeze(indices).to('cpu')
    final_output = smu.get_person_proposal_from_network_output(outputs, subgraph, indices, nodes_camera,
                                                               scenario.jsons_for_head, 0.5)
    batched_input, raw_inputs = [], []
    for person in final_output:                                          # :243-275
        raw_input = {}
        for camera in cfg.used_pe_names:
            if person[camera] is not None:
                raw_input[camera] = [json.dumps([scenario.jsons_for_head[person[camera]]])]
        inputs = ds.PoseEstimatorDataset(raw_input, list(range(cfg.n_cameras)), list(range(18)), save=False)
        if inputs.__len__() == 0:
            continue
        batched_input.append(inputs[0][0].reshape([1, inputs[0][0].size()[0]]).to(device))
        raw_inputs.append(raw_input)
    results = None
    if batched_input:
        input_all = torch.cat(batched_input, dim=0)
        output_all = mlp(input_all.to(device))
        results = (output_all * 10.).to('cpu').numpy()
    return dict(scenario=scenario, subgraph=subgraph, outputs=outputs.cpu().numpy(), indices=indices.numpy(),
                nodes_camera=nodes_camera, final_output=final_output, mlp_in=[b.cpu().numpy()[0] for b in batched_input],
                results=results, raw_inputs=raw_inputs)


@pytest.mark.parametrize('config', ['panoptic', 'arp3', 'arp6', 'arp_robot2', 'pansub'])
def test_driver_loop_against_reference_goldens(config):
    cfg, npz, meta = helpers.load_golden(config)
    mods = dropin_env.activate(cfg)
    gat_state, mlp_state = helpers.golden_weights(config)
    device = torch.device('cuda')
    assert len(mods['graph_generator'].HumanGraphFromView.get_all_features('3')) == cfg.n_features_sm
    model = mods['gat2'].GAT2(None, 5, cfg.n_features_sm, 1, [40, 40, 40, 30], [10, 10, 8, 5], torch.nn.LeakyReLU(),
                              torch.nn.Sigmoid(), 0., 0., 0.15, False, bias=True)
    model.load_state_dict(gat_state)
    model = model.to(device)
    mlp = mods['mlp'].PoseEstimatorMLP(input_dimensions=meta['mlp_in_dim'], output_dimensions=54)   # show_results_from_model.py:123
    mlp.load_state_dict(mlp_state)
    mlp = mlp.to(device)
    pu = mods['pose_estimator_utils']
    names = cfg.used_sm_names
    n_checked = 0
    explained = []
    tabs = O.CameraTables(cfg)
    for tag in meta['cases']:
        out = run_frame(mods, cfg, model, mlp, meta['frames'][tag])
        if tag in meta['no_graph']:
            assert out is None
            continue
        g = out['subgraph']
        src, dst = [x.tolist() for x in g.edges()]
        assert src == npz[tag + '/src'].tolist() and dst == npz[tag + '/dst'].tolist()
        assert g.number_of_nodes() == int(npz[tag + '/n_nodes'])
        assert g.nodes().tolist() == list(range(g.number_of_nodes()))
        assert np.array_equal(g.ndata['h'].cpu().numpy(), npz[tag + '/feats'])
        assert np.array_equal(out['indices'].reshape(-1), npz[tag + '/indices'].reshape(-1))   # 0-dim when M == 1
        want_cam = ['' if c < 0 else names[c] for c in npz[tag + '/nodes_camera']]
        assert out['nodes_camera'] == want_cam
        assert tuple(out['scenario'].labels[0].shape) == (out['indices'].size, 1)
        assert out['scenario'].labels[0].dtype == torch.float64
        assert g.edata['rel_type'].tolist() == npz[tag + '/rel_type'].tolist()
        assert np.array_equal(g.edata['norm'].cpu().numpy(), npz[tag + '/norm'])
        idx = npz[tag + '/indices'].reshape(-1)
        ref = npz[tag + '/scores']
        rel = np.abs(out['outputs'][idx] - ref[idx]) / np.abs(ref[idx])
        assert rel.max() <= 1e-4, (tag, rel.max())
        want = npz[tag + '/proposals']
        got = np.array([[-1 if p[c] is None else p[c] for c in names] for p in out['final_output']], dtype=np.int32).reshape(-1, len(names))
        if not np.array_equal(got, want):
            # accepted only with a full attribution to a score gap below the tolerance; raises otherwise
            f = meta['frames'][tag]
            og = O.build_graph({c: f[c] for c in f if json.loads(f[c][0])}, tabs)
            explained.append((tag, OC.explain_assignment_mismatch(ref, out['outputs'], og, cfg, got)))
            continue
        n_checked += 1
        for p in range(len(want)):
            assert np.abs(out['mlp_in'][p] - npz[tag + '/mlp_in'][p]).max() <= 1e-6, (tag, p)
        if len(want):
            ref_j = npz[tag + '/mlp_out'] * np.float32(10.)
            assert np.abs(out['results'] - ref_j).max() <= 0.5e-3, tag
        # triangulation baseline, fed like test/metrics_from_triangulation.py:237-249
        cam_matrix = {c: cfg.K32(i) for i, c in enumerate(cfg.camera_names)}
        dist = {c: cfg.dist64(i) for i, c in enumerate(cfg.camera_names)}
        proj = {c: cfg.P64(i) for i, c in enumerate(cfg.camera_names)}
        for p, raw_input in enumerate(out['raw_inputs']):
            points_2D = {}
            for camera in raw_input:
                sk = json.loads(raw_input[camera][0])[0]
                for j, values in sk.items():
                    points_2D.setdefault(j, {})[camera] = [values[1], values[2]]
            res = pu.triangulate(points_2D, cam_matrix, dist, proj, cfg.median_axis)
            mask = npz[tag + '/tri_mask'][p]
            assert sorted(int(j) for j in res) == [j for j in range(18) if mask[j]]
            for j, X in res.items():
                assert X.shape == (3, 1)
                assert np.abs(X[:, 0] - npz[tag + '/tri'][p][int(j)]).max() <= 1e-7
    print('drop-in driver loop %s: %d frames equal the reference, explained near-ties: %s' % (config, n_checked, explained))
    assert n_checked > 0


def test_proposals_accept_python_lists_and_custom_threshold():
    cfg, npz, meta = helpers.load_golden('panoptic')
    mods = dropin_env.activate(cfg)
    tag = 'p4a'
    frame = {c: v for c, v in meta['frames'][tag].items() if json.loads(v[0])}
    scenario = mods['graph_generator'].MergedMultipleHumansDataset(frame, mode='test', alt='3', debug=True, verbose=False)
    g = scenario.graphs[0]
    scores = npz[tag + '/scores']
    out = mods['skeleton_matching_utils'].get_person_proposal_from_network_output(
        [float(s) for s in scores], g, None, scenario.data['nodes_camera'][0], scenario.jsons_for_head)
    names = cfg.used_sm_names
    got = np.array([[-1 if p[c] is None else p[c] for c in names] for p in out], dtype=np.int32).reshape(-1, len(names))
    assert np.array_equal(got, npz[tag + '/proposals'])
    none = mods['skeleton_matching_utils'].get_person_proposal_from_network_output(
        torch.from_numpy(scores), g, None, scenario.data['nodes_camera'][0], None, 1.5)
    assert none == []
    hg = mods['graph_generator'].HumanGraphFromView(scenario.jsons_for_head[0], scenario.data['nodes_camera'][0][0], '3')
    assert hg.n_nodes == 1 and hg.src_nodes == [0] and hg.num_joints == len(scenario.jsons_for_head[0])
    assert np.array_equal(hg.features.cpu().numpy()[0], npz[tag + '/feats'][0])


def test_live_frames_fall_back_when_the_submission_cannot_answer():
    """The whole-frame submissions behind the drop-in modules (live.py) must never answer a call they do not match: a graph
    whose submission was overtaken by the next frame of the same shape, another threshold, scores that are not the
    submission's, a person that is not a proposal, changed weights - each takes the eager path and still gives the
    reference's result."""
    config = 'panoptic'
    cfg, npz, meta = helpers.load_golden(config)
    mods = dropin_env.activate(cfg)
    rt = mods['rt']
    gat_state, mlp_state = helpers.golden_weights(config)
    device = torch.device('cuda')
    model = mods['gat2'].GAT2(None, 5, cfg.n_features_sm, 1, [40, 40, 40, 30], [10, 10, 8, 5], torch.nn.LeakyReLU(),
                              torch.nn.Sigmoid(), 0., 0., 0.15, False, bias=True)
    model.load_state_dict(gat_state)
    model = model.to(device)
    mlp = mods['mlp'].PoseEstimatorMLP(input_dimensions=meta['mlp_in_dim'], output_dimensions=54)
    mlp.load_state_dict(mlp_state)
    mlp = mlp.to(device)
    names = cfg.used_sm_names
    arr = lambda fo: np.array([[-1 if p[c] is None else p[c] for c in names] for p in fo], dtype=np.int32).reshape(-1, len(names))
    for tag in ('p4a', 'p4b'):                                   # first frames: eager, they tell the runtime which models run
        run_frame(mods, cfg, model, mlp, meta['frames'][tag])
    assert rt.live() is not None
    gg, smu = mods['graph_generator'], mods['skeleton_matching_utils']
    fa = {c: [v[0], v[1]] for c, v in meta['frames']['p4a'].items()}
    fb = {c: [v[0], v[1]] for c, v in meta['frames']['p4b'].items()}          # same shape as p4a: reuses its graphs
    sa = gg.MergedMultipleHumansDataset(fa, mode='test', alt='3', debug=True, verbose=False)
    ga = sa.graphs[0]
    assert hasattr(ga, '_live') and ga._live.fresh()
    feats_a = ga.ndata['h']
    sb = gg.MergedMultipleHumansDataset(fb, mode='test', alt='3', debug=True, verbose=False)
    assert not ga._live.fresh() and sb.graphs[0]._live.fresh()
    # the overtaken graph still gives its own frame's features, scores and proposals
    assert np.array_equal(ga.ndata['h'].cpu().numpy(), npz['p4a/feats'])
    out_a = torch.squeeze(model(ga.ndata['h'].float(), ga))
    idx = npz['p4a/indices']
    ref = npz['p4a/scores']
    assert (np.abs(out_a.cpu().numpy()[idx] - ref[idx]) / np.abs(ref[idx])).max() <= 1e-4
    pa = smu.get_person_proposal_from_network_output(out_a, ga, None, sa.data['nodes_camera'][0], sa.jsons_for_head, 0.5)
    assert np.array_equal(arr(pa), npz['p4a/proposals'])
    # the fresh one answers from its submission; another threshold or foreign scores do not
    gb = sb.graphs[0]
    out_b = torch.squeeze(model(gb.ndata['h'].float(), gb))
    pb_ = smu.get_person_proposal_from_network_output(out_b, gb, None, sb.data['nodes_camera'][0], sb.jsons_for_head, 0.5)
    assert np.array_equal(arr(pb_), npz['p4b/proposals'])
    assert smu.get_person_proposal_from_network_output(out_b, gb, None, sb.data['nodes_camera'][0], None, 1.5) == []
    fuzz = torch.from_numpy(npz['fuzz/1/scores']).cuda() if meta['fuzz_tags'][1] == 'p4b' else None
    other = torch.from_numpy(npz['p4b/scores']).cuda()
    po = smu.get_person_proposal_from_network_output(other, gb, None, sb.data['nodes_camera'][0], None, 0.5)
    assert np.array_equal(arr(po), npz['p4b/proposals'])
    # a person that is not one of the proposals (heads 0 and 5 of two cameras, whatever the clustering said): encoded eagerly
    ds = mods['pose_estimator_dataset_from_json']
    h0, h1 = 0, next(h for h, c in enumerate(sb.data['nodes_camera'][0]) if c != sb.data['nodes_camera'][0][0])
    raw = {sb.data['nodes_camera'][0][h0]: [json.dumps([sb.jsons_for_head[h0]])],
           sb.data['nodes_camera'][0][h1]: [json.dumps([sb.jsons_for_head[h1]])]}
    row = ds.PoseEstimatorDataset(raw, list(range(cfg.n_cameras)), list(range(18)), save=False)[0][0]
    import importlib
    O = importlib.import_module('oracle.pose_oracle')
    want = O.encode_person({c: json.loads(v[0])[0] for c, v in raw.items()}, O.CameraTables(cfg))
    assert np.abs(row.cpu().numpy() - want).max() <= 1e-6
    # changed weights: the next frame is eager again and correct for the NEW weights
    with torch.no_grad():
        model.layers[4].fc2.bias.add_(0.25)
    sc = gg.MergedMultipleHumansDataset(fa, mode='test', alt='3', debug=True, verbose=False)
    out_c = torch.squeeze(model(sc.graphs[0].ndata['h'].float(), sc.graphs[0])).cpu().numpy()
    gat_np = helpers.np_state(gat_state)
    gat_np['layers.4.fc2.bias'] = gat_np['layers.4.fc2.bias'] + np.float32(0.25)
    ref_c = O.gat_forward(gat_np, npz['p4a/feats'], npz['p4a/src'], npz['p4a/dst'])
    assert (np.abs(out_c[idx] - ref_c[idx]) / np.abs(ref_c[idx])).max() <= 1e-4
