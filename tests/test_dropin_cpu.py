"""CPU: the drop-in modules import under the reference's module names, expose its API surface, and load
state dicts with the reference's key names. (Running them needs the GPU: tests/test_dropin_gpu.py.)"""
import inspect

import pytest
import torch

import dropin_env
import helpers


@pytest.fixture(scope='module')
def mods():
    cfg, npz, meta = helpers.load_golden('panoptic')
    return dropin_env.activate(cfg)


def test_reference_names_and_signatures(mods):
    gg = mods['graph_generator']
    assert list(inspect.signature(gg.MergedMultipleHumansDataset.__init__).parameters)[1:] == [
        'paths', 'probabilities', 'limit', 'alt', 'mode', 'force_reload', 'verbose', 'debug', 'raw_dir']
    assert gg.graphData._fields == ('src_nodes', 'dst_nodes', 'n_nodes', 'features', 'edge_types', 'edge_norms')
    assert len(gg.HumanGraphFromView.get_all_features('3')) == 902           # 2 + 180 * 5 views
    assert len(gg.HumanGraphFromView.get_all_features()) == 2 + 18 + 5 + 4 + 1   # alternative '1' vocabulary
    assert gg.HumanGraphFromView.get_all_features('3')[:2] == ['head', 'edge_node']
    assert gg.HumanGraphFromView.get_rels('3') == ['h_h', 'link', 'link_link']
    assert list(inspect.signature(mods['gat2'].GAT2.__init__).parameters)[1:] == [
        'g', 'num_layers', 'in_dim', 'num_classes', 'num_hidden', 'heads', 'activation', 'final_activation', 'feat_drop',
        'attn_drop', 'alpha', 'residual', 'bias']
    assert list(inspect.signature(mods['skeleton_matching_utils'].get_person_proposal_from_network_output).parameters) == [
        'outputs', 'subgraph', 'indices', 'nodes_camera', 'jsons_for_head', 'CLASSIFICATION_THRESHOLD']
    assert list(inspect.signature(mods['pose_estimator_utils'].triangulate).parameters) == [
        'points_2D', 'camera_matrices', 'distortion_coefficients', 'projection_matrices', 'median_chek_axis']
    assert list(inspect.signature(mods['pose_estimator_dataset_from_json'].PoseEstimatorDataset.__init__).parameters)[1:] == [
        'input_data', 'cameras', 'joint_list', 'transform', 'data_augmentation', 'reload', 'save', 'device']
    for name in ('camera_matrix', 'from_homogeneous', 'from_homogeneous2', 'get_distortion_coefficients', 'apply_distortion', 'device'):
        assert hasattr(mods['pose_estimator_utils'], name)


def test_state_dict_keys_match_the_reference(mods):
    gat_state, mlp_state = helpers.golden_weights('panoptic')
    model = mods['gat2'].GAT2(None, 5, 902, 1, [40, 40, 40, 30], [10, 10, 8, 5], torch.nn.LeakyReLU(), torch.nn.Sigmoid(),
                              0., 0., 0.15, False, bias=True)
    assert set(model.state_dict().keys()) == set(gat_state.keys())
    model.load_state_dict(gat_state)
    mlp = mods['mlp'].PoseEstimatorMLP(input_dimensions=5 * 18 * 14, output_dimensions=54)
    assert set(mlp.state_dict().keys()) == set(mlp_state.keys())
    mlp.load_state_dict(mlp_state)
    model.set_g('g')
    assert all(layer.g == 'g' for layer in model.layers)


def test_small_helpers(mods):
    pu = mods['pose_estimator_utils']
    K = pu.camera_matrix(1, use_cuda=False)
    cfg, _, _ = helpers.load_golden('panoptic')
    assert K.dtype == torch.float32 and torch.equal(K, torch.from_numpy(cfg.K32(1)))
    v = torch.tensor([[2.0, 4.0], [6.0, 8.0], [2.0, 4.0]])
    assert torch.equal(pu.from_homogeneous(v), torch.tensor([[1.0, 1.0], [3.0, 2.0]]))
    assert torch.equal(pu.from_homogeneous2(v)[2], torch.ones(2))
    kd = torch.tensor([0.1, 0.01, 0.001])
    d = pu.apply_distortion(kd, torch.tensor([[0.5], [0.25], [1.0]]))
    r2 = 0.5 ** 2 + 0.25 ** 2
    assert abs(float(d[0, 0]) - 0.5 * (1 + 0.1 * r2 + 0.01 * r2 ** 2 + 0.001 * r2 ** 3)) < 1e-6 and float(d[2, 0]) == 1.0
    ds = mods['pose_estimator_dataset_from_json']
    assert ds.get_skeleton_indices({'a': ['[{"1": [1,0,0,1,1]}, {"1": [1,0,0,1,1], "2": [2,0,0,1,1]}]']}) == {'a': 1}


def test_out_of_scope_modes_fail_loudly(mods):
    gg = mods['graph_generator']
    with pytest.raises(NotImplementedError):
        gg.MergedMultipleHumansDataset({}, mode='test', alt='1')
    with pytest.raises(NotImplementedError):
        gg.MergedMultipleHumansDataset({}, mode='train', alt='2')
    with pytest.raises(SystemExit):
        gg.MergedMultipleHumansDataset({}, mode='test', alt=None)
    with pytest.raises(NotImplementedError):
        mods['pose_estimator_dataset_from_json'].PoseEstimatorDataset(['x.json'], [0], [0])
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):                       # no CPU fallback
            gg.MergedMultipleHumansDataset({'trackera': ['[{"1": [1, 5.0, 5.0, 1, 1]}]', 0.0]}, mode='test', alt='3')


def test_dgl_drop_in_offers_batch_only(mods):
    """`import dgl` resolves to the drop-in's module (train_skeleton_matching.py:11,80; sm_metrics_without_gt.py:6,60):
    `batch` for the B200 graphs, a clear error for everything else."""
    import importlib
    import sys
    sys.modules.pop('dgl', None)
    dgl = importlib.import_module('dgl')
    assert dgl.__file__.startswith(dropin_env.SHADOW)
    assert list(inspect.signature(dgl.batch).parameters)[0] == 'graphs'
    with pytest.raises(ValueError):
        dgl.batch([])
    with pytest.raises(TypeError):
        dgl.batch([object()])
    with pytest.raises(AttributeError):
        dgl.graph
    sys.modules.pop('dgl', None)
