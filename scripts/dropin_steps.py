"""Wall time of each step of the reference driver's per-frame loop body (test/metrics_from_model.py:178-300) on the
drop-in modules: where a live frame's milliseconds go, split into what the driver script itself spends (JSON re-encoding
per camera and per person, its own host<->device copies) and what the drop-in modules spend.

    python scripts/dropin_steps.py            (B200POSE_DROPIN_LIVE=0 disables the whole-frame submissions)
"""
import collections, contextlib, io, json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (REPO, os.path.join(REPO, 'tests')):
    if p_ not in sys.path:
        sys.path.insert(0, p_)

DRIVER_OWN = ('driver: json.loads + json.dumps per camera', 'driver: jsons_for_head + json.dumps per person', 'driver: row.to(device)',
              'driver: x10 + .to(cpu) per person', 'driver: .to(device) of graph / indices / features')


def measure(cfg, frames, gat, mlp_state, warm=8):
    """Returns dict(frames_per_s, ms_per_frame, driver_own_ms, dropin_ms, steps={name: ms per frame})."""
    import torch
    import dropin_env
    T = collections.OrderedDict()
    torch.set_grad_enabled(False)                                # test/metrics_from_model.py:54
    with contextlib.redirect_stdout(io.StringIO()):             # the reference's modules print while they work
        mods = dropin_env.activate(cfg)
        dev = torch.device('cuda')
        model = mods['gat2'].GAT2(None, 5, cfg.n_features_sm, 1, [40, 40, 40, 30], [10, 10, 8, 5], torch.nn.LeakyReLU(), torch.nn.Sigmoid(),
                                  0., 0., 0.15, False, bias=True)
        model.load_state_dict(gat); model = model.to(dev)
        mlp = mods['mlp'].PoseEstimatorMLP(input_dimensions=cfg.mlp_in, output_dimensions=54)
        mlp.load_state_dict(mlp_state); mlp = mlp.to(dev)
        gg, smu, ds = mods['graph_generator'], mods['skeleton_matching_utils'], mods['pose_estimator_dataset_from_json']

        def tick(name, t0):
            t = time.perf_counter(); T.setdefault(name, []).append(t - t0); return t

        def one(frame):
            t = time.perf_counter()
            processed_input = {}
            for cam in frame:                                        # metrics_from_model.py:182-191
                cam_data = json.loads(frame[cam][0])
                if cam_data:
                    processed_input[cam] = [json.dumps(cam_data), frame[cam][1]]
            t = tick('driver: json.loads + json.dumps per camera', t)
            scenario = gg.MergedMultipleHumansDataset(processed_input, mode='test', limit=10000, debug=True, alt='3', verbose=False)
            t = tick('MergedMultipleHumansDataset', t)
            if len(scenario.graphs) == 0:
                return
            subgraph = scenario.graphs[0].to(dev)
            indices = scenario.data['edge_nodes_indices'][0].to(dev)
            nodes_camera = scenario.data['nodes_camera'][0]
            feats = subgraph.ndata['h'].to(dev)
            t = tick('driver: .to(device) of graph / indices / features', t)
            model.g = subgraph
            for layer in model.layers:
                layer.g = subgraph
            outputs = torch.squeeze(model(feats.float(), subgraph))
            t = tick('model(feats, subgraph)', t)
            indices = torch.squeeze(indices).to('cpu')
            t = tick('indices.to(cpu)  [waits for the GPU]', t)
            final_output = smu.get_person_proposal_from_network_output(outputs, subgraph, indices, nodes_camera, scenario.jsons_for_head, 0.5)
            t = tick('get_person_proposal_from_network_output', t)
            batched = []
            for person in final_output:                              # :243-275
                raw_input = {}
                for camera in cfg.used_pe_names:
                    if person[camera] is not None:
                        raw_input[camera] = [json.dumps([scenario.jsons_for_head[person[camera]]])]
                t = tick('driver: jsons_for_head + json.dumps per person', t)
                inputs = ds.PoseEstimatorDataset(raw_input, list(range(cfg.n_cameras)), list(range(18)), save=False)
                t = tick('PoseEstimatorDataset', t)
                batched.append(inputs[0][0].reshape([1, inputs[0][0].size()[0]]).to(dev))
                t = tick('driver: row.to(device)', t)
            if batched:
                out = mlp(torch.cat(batched, dim=0).to(dev))
                t = tick('mlp(input_all)', t)
                [(torch.squeeze(out[i]) * 10.).to('cpu') for i in range(out.shape[0])]      # :281-283
                t = tick('driver: x10 + .to(cpu) per person', t)

        for f in frames[:warm]:
            one(f)
        T.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for f in frames:
            one(f)
        torch.cuda.synchronize()
        total = time.perf_counter() - t0
    n = len(frames)
    steps = {k: 1e3 * sum(v) / n for k, v in T.items()}
    own = sum(v for k, v in steps.items() if k in DRIVER_OWN)
    return dict(frames_per_s=n / total, ms_per_frame=1e3 * total / n, driver_own_ms=own, dropin_ms=1e3 * total / n - own, steps=steps)


if __name__ == '__main__':
    import importlib
    import bench
    synth = importlib.import_module('3d_multi_pose_estimator_b200.synth')
    pkg = importlib.import_module('3d_multi_pose_estimator_b200')
    cfg = pkg.CameraConfig.from_npz(os.path.join(REPO, 'tests', 'golden', 'cameras_panoptic.npz'))
    frames = [synth.make_frame(cfg, i, 4) for i in range(64)]
    gat, mlp_state = bench.load_weights('panoptic', cfg)
    r = measure(cfg, frames, gat, mlp_state)
    print("%d frames, %.3f ms per frame (%.0f frames/s): the driver script's own work %.3f ms, the drop-in modules %.3f ms; live path %s" % (
        len(frames), r['ms_per_frame'], r['frames_per_s'], r['driver_own_ms'], r['dropin_ms'], os.environ.get('B200POSE_DROPIN_LIVE', '1')))
    for k, v in r['steps'].items():
        print('  %-52s %7.3f ms per frame' % (k, v))
