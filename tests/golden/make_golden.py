"""Regenerates tests/golden/*.npz by running the UNMODIFIED reference (under /root/reference, CPU,
with the two import shims of oracle/shims) on seeded synthetic frames.

    python tests/golden/make_golden.py [panoptic|arp3|ring10 ...]

Only runs in the build container (needs /root/reference, cv2, networkx). The fixtures it writes are what
travels to the GPU box. The per-frame glue below restates the call sequence of the reference's
test/metrics_from_model.py:178-300 and test/metrics_from_triangulation.py:234-249; every number stored
is produced by the reference's own functions.

One sub-process per configuration: the reference keeps its configuration in module globals.
"""
import json
import os
import pickle
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

CALIB_GAIN = 1.0    # gain on the last layer (see calibrate()); 1.0 = bias shift only
GAT_SEED = 0
MLP_SEED = 1


def _cases(config):
    """(tag, n_persons, seed, synth kwargs)"""
    if config == 'panoptic':
        return [('p3', 3, 0, {}), ('p4a', 4, 0, {}), ('p4b', 4, 1, {}), ('p4c', 4, 2, {}),
                ('rag1', 4, 10, dict(drop_joint_p=0.25, drop_view_p=0.25, rand_conf=True, keep_empty=True)),
                ('rag2', 5, 11, dict(drop_joint_p=0.5, drop_view_p=0.4, rand_conf=True, keep_empty=True)),
                ('order', 3, 12, dict(camera_order=[3, 0, 4])),
                ('m1', 1, 13, dict(camera_order=[2, 1])),
                ('onecam', 3, 14, dict(camera_order=[1])),
                ('p6', 6, 15, dict(drop_joint_p=0.1))]
    if config == 'arp3':
        return [('p8a', 8, 0, {}), ('p8b', 8, 1, {}),
                ('rag', 6, 2, dict(drop_joint_p=0.3, drop_view_p=0.3, rand_conf=True, keep_empty=True))]
    if config == 'ring10':
        return [('p3', 3, 0, {}), ('rag', 4, 1, dict(drop_joint_p=0.3, drop_view_p=0.3, rand_conf=True))]
    if config == 'arp6':
        # the reference's shipped ARPLAB configuration (parameters.py:79-123): four room cameras + the narrow-baseline
        # orinbot_l / orinbot_r stereo pair (0.12 m). 'stereo*': persons seen by the pair only, exact and with pixel noise
        # (the ill-conditioned DLT of SURVEY.md 7-7); 'mixed': one room camera + the pair.
        return [('p4', 4, 0, {}), ('p3n', 3, 1, dict(pixel_noise=1.5)),
                ('rag', 5, 2, dict(drop_joint_p=0.3, drop_view_p=0.3, rand_conf=True, keep_empty=True, pixel_noise=0.7)),
                ('stereo', 3, 3, dict(camera_order=[4, 5])), ('stereo_n', 4, 4, dict(camera_order=[5, 4], pixel_noise=1.0)),
                ('mixed', 3, 5, dict(camera_order=[4, 1, 5], pixel_noise=0.5, drop_joint_p=0.2))]
    if config == 'arp_robot2':
        # parameters.py:110-112: models using only the robot cameras. Frames still carry every camera of the rig.
        return [('p3', 3, 0, dict(all_cameras=True)), ('p4n', 4, 1, dict(all_cameras=True, pixel_noise=1.0)),
                ('rag', 5, 2, dict(all_cameras=True, drop_joint_p=0.3, drop_view_p=0.2, rand_conf=True, keep_empty=True)),
                ('only', 2, 3, dict(camera_order=[5, 4])), ('unused_only', 3, 4, dict(camera_order=[0, 2, 4]))]
    if config == 'pansub':
        # Panoptic with used_cameras_skeleton_matching = 4 of the 5 cameras in a permuted order and used_cameras = 3 of
        # those; frames carry all five cameras (graph_generator.py:583-584 skips the unused one).
        return [('p3', 3, 0, dict(all_cameras=True)), ('p4', 4, 1, dict(all_cameras=True)),
                ('rag', 5, 2, dict(all_cameras=True, drop_joint_p=0.3, drop_view_p=0.3, rand_conf=True, keep_empty=True)),
                ('order', 3, 3, dict(camera_order=[4, 0, 1, 3])), ('pe_absent', 3, 4, dict(camera_order=[2, 0, 3]))]
    raise ValueError(config)


def _activate(config):
    from oracle import ref_env
    return ref_env.activate_config(config)


def build_reference_models(parameters, n_feats):
    """Random-init weights of the reference architectures (train_skeleton_matching.py:40-56,148-149;
    metrics_from_model.py:90-100), seeded so the GPU box can rebuild the same tensors."""
    import torch
    from gat2 import GAT2
    from mlp import PoseEstimatorMLP
    torch.manual_seed(GAT_SEED)
    gat = GAT2(None, 5, n_feats, 1, [40, 40, 40, 30], [10, 10, 8, 5], torch.nn.LeakyReLU(), torch.nn.Sigmoid(),
               0., 0., 0.15, False, bias=True)
    torch.manual_seed(MLP_SEED)
    # metrics_from_model.py:91 sizes the MLP with len(parameters.cameras), show_results_from_model.py:123 with
    # len(parameters.used_cameras); the encoder emits len(used_cameras) blocks (dataset.py:129), so only the latter
    # runs when used_cameras is a proper subset.
    mlp = PoseEstimatorMLP(input_dimensions=len(parameters.used_cameras) * len(parameters.joint_list) * parameters.numbers_per_joint,
                           output_dimensions=54)
    return gat.eval(), mlp.eval()


def checksum(state_dict):
    out = {}
    for k, v in state_dict.items():
        a = v.detach().double().numpy().ravel()
        out[k] = [float(a.sum()), float(np.abs(a).sum()), float(a[0]), float(a[-1])]
    return out


def reference_frame(frame, parameters, gat, mlp, mods):
    """One frame through the reference path. Returns dict of numpy arrays (or None if no graph)."""
    import torch
    MergedMultipleHumansDataset = mods['graph_generator'].MergedMultipleHumansDataset
    get_pp = mods['skeleton_matching_utils'].get_person_proposal_from_network_output
    PoseEstimatorDataset = mods['dataset'].PoseEstimatorDataset
    processed_input = {}
    for cam in frame:                                     # metrics_from_model.py:182-191
        data = json.loads(frame[cam][0])
        if data:
            processed_input[cam] = [json.dumps(list(data)), frame[cam][1]]
    scenario = MergedMultipleHumansDataset(processed_input, mode='test', limit=10000, debug=True,
                                           alt=parameters.graph_alternative, verbose=False)
    if len(scenario.graphs) == 0:
        return None
    g = scenario.graphs[0]
    indices = scenario.data['edge_nodes_indices'][0]
    nodes_camera = scenario.data['nodes_camera'][0]
    feats = g.ndata['h']
    layer_out = []
    hooks = [l.register_forward_hook(lambda m, i, o: layer_out.append(o.detach().clone())) for l in gat.layers]
    outputs = torch.squeeze(gat(feats.float(), g))
    for h in hooks:
        h.remove()
    idx_sq = torch.squeeze(indices).to('cpu')
    final_output = get_pp(outputs, g, idx_sq, nodes_camera, scenario.jsons_for_head, 0.5)
    src, dst = [x.numpy().astype(np.int32) for x in g.edges()]
    cam_names = list(parameters.used_cameras_skeleton_matching)
    rec = dict(src=src, dst=dst, n_nodes=np.int32(g.number_of_nodes()), feats=feats.numpy(),
               indices=indices.numpy().ravel().astype(np.int64),
               nodes_camera=np.array([cam_names.index(c) if c != '' else -1 for c in nodes_camera], dtype=np.int32),
               skeleton_index=np.array([scenario.skeleton_index[h] for h in sorted(scenario.skeleton_index)], dtype=np.int32),
               scores=outputs.numpy().reshape(-1).astype(np.float32),
               label_shape=np.array(scenario.labels[0].shape),
               rel_type=g.edata['rel_type'].numpy().astype(np.int64), norm=g.edata['norm'].numpy())
    for l, o in enumerate(layer_out):
        rec['gat_l%d' % l] = o.numpy()
    rec['proposals'] = proposals_to_array(final_output, cam_names)
    # ---- stage 3 (metrics_from_model.py:243-280) + triangulation baseline (metrics_from_triangulation.py:234-249)
    stage3(rec, '', final_output, scenario, parameters, mlp, mods)
    # the same stage on a FORCED assignment (k-th head of every camera = person k), so the encoder / DLT tiers are
    # covered on every golden frame whatever the random-init scores make of the clustering
    per_cam = {}
    for h, c in enumerate(nodes_camera):
        if c != '':
            per_cam.setdefault(c, []).append(h)
    forced = []
    for k in range(max(len(v) for v in per_cam.values())):
        person = {c: (per_cam[c][k] if c in per_cam and k < len(per_cam[c]) else None) for c in cam_names}
        forced.append(person)
    rec['forced_proposals'] = proposals_to_array(forced, cam_names)
    stage3(rec, 'forced_', forced, scenario, parameters, mlp, mods)
    return rec


def stage3(rec, prefix, final_output, scenario, parameters, mlp, mods):
    import torch
    PoseEstimatorDataset = mods['dataset'].PoseEstimatorDataset
    mlp_in, tri, tri_mask, enc_ok = [], [], [], []
    dsm = mods['dataset']
    for person in final_output:
        raw_input = {}
        points_2D = {}
        for camera in parameters.used_cameras:
            if person.get(camera) is not None:
                sk = scenario.jsons_for_head[person[camera]]
                raw_input[camera] = [json.dumps([sk])]
                for j, values in sk.items():
                    points_2D.setdefault(j, {})[camera] = [values[1], values[2]]
        try:
            ds = PoseEstimatorDataset(raw_input, parameters.cameras, parameters.joint_list, save=False)
            mlp_in.append(ds[0][0].numpy()); enc_ok.append(1)
        except RuntimeError:          # torch.stack([]) when the row is (almost) empty, dataset.py:287-298
            mlp_in.append(np.zeros(len(parameters.used_cameras) * 252, np.float32)); enc_ok.append(0)
        r = mods['utils'].triangulate(points_2D, dsm.camera_matrices, dsm.distortion_coefficients,
                                      dsm.projection_matrices, parameters.axes_3D['Y'][0])
        t = np.zeros((18, 3)); m = np.zeros(18, dtype=np.uint8)
        for j, v in r.items():
            t[int(j)] = np.asarray(v).ravel(); m[int(j)] = 1
        tri.append(t); tri_mask.append(m)
    if mlp_in:
        x = torch.from_numpy(np.stack(mlp_in))
        rec[prefix + 'mlp_in'] = x.numpy()
        rec[prefix + 'mlp_out'] = mlp(x).numpy()
        rec[prefix + 'enc_ok'] = np.array(enc_ok, np.uint8)
        rec[prefix + 'tri'] = np.stack(tri); rec[prefix + 'tri_mask'] = np.stack(tri_mask)


def proposals_to_array(final_output, cam_names):
    arr = np.full((len(final_output), len(cam_names)), -1, dtype=np.int32)
    for p, person in enumerate(final_output):
        for c, name in enumerate(cam_names):
            if person[name] is not None:
                arr[p, c] = person[name]
    return arr


def calibrate(gat, frames, parameters, mods):
    """Random-init scores sit in a narrow band on one side of 0.5, which makes the
    clustering trivially all-pass/all-fail. Shift the last layer's fc2 bias so the median edge-node
    logit is 0 on the given frames (the last layer's output is sum_u alpha_u*(w.h2_u + b) with
    sum alpha = 1, so this is an exact shift of every logit). A gain > 1 would spread the scores but
    multiplies fp32 rounding noise and score gaps alike (measured: x24 gain puts two fp32
    implementations 1.1e-4 apart), so it buys nothing and is left at 1."""
    import torch
    MergedMultipleHumansDataset = mods['graph_generator'].MergedMultipleHumansDataset
    sig = gat.final_activation
    gat.final_activation = None
    last = gat.layers[-1]

    def logits():
        out = []
        for frame in frames:
            pi = {c: [frame[c][0], frame[c][1]] for c in frame if json.loads(frame[c][0])}
            sc = MergedMultipleHumansDataset(pi, mode='test', limit=10000, debug=True, alt=parameters.graph_alternative, verbose=False)
            if not sc.graphs:
                continue
            o = torch.squeeze(gat(sc.graphs[0].ndata['h'].float(), sc.graphs[0]))
            out.append(o[sc.data['edge_nodes_indices'][0].ravel()])
        return torch.cat(out)
    with torch.no_grad():
        z = logits()
        gain = float(CALIB_GAIN)
        last.fc2.weight.mul_(gain); last.fc2.bias.mul_(gain)
        z = logits()
        shift = float(z.median())
        last.fc2.bias.sub_(shift)
        z = logits()
    gat.final_activation = sig
    return gain, shift, z.numpy()


def cluster_fuzz(frames_graphs, parameters, mods, rng, n_cases):
    """Clustering on fuzzed score vectors (ties, all-pass, all-fail, random): SURVEY.md App. B."""
    get_pp = mods['skeleton_matching_utils'].get_person_proposal_from_network_output
    cam_names = list(parameters.used_cameras_skeleton_matching)
    out = []
    for case in range(n_cases):
        tag, g, indices, nodes_camera = frames_graphs[case % len(frames_graphs)]
        n = g.number_of_nodes()
        mode = case % 6
        if mode == 0:
            s = rng.uniform(0, 1, n)
        elif mode == 1:
            s = np.round(rng.uniform(0, 1, n), 1)            # many exact ties
        elif mode == 2:
            s = rng.uniform(0.5, 1, n)                       # all pass
        elif mode == 3:
            s = rng.uniform(0.3, 0.7, n)
        elif mode == 4:
            s = np.where(rng.random(n) < 0.3, 0.9, rng.uniform(0, 0.6, n))
        else:
            s = rng.beta(0.3, 0.3, n)                        # bimodal, like a trained matcher
        s = s.astype(np.float32)
        import torch
        res = get_pp(torch.from_numpy(s), g, indices, nodes_camera, None, 0.5)
        out.append((tag, s, proposals_to_array(res, cam_names)))
    return out


def run_config(config):
    import torch
    torch.set_grad_enabled(False)
    parameters = _activate(config)
    import gat2, graph_generator, mlp as mlp_mod, pose_estimator_dataset_from_json as dataset  # noqa: E401
    import pose_estimator_utils, skeleton_matching_utils
    mods = dict(graph_generator=graph_generator, dataset=dataset, utils=pose_estimator_utils,
                skeleton_matching_utils=skeleton_matching_utils)
    b200 = __import__('importlib').import_module('3d_multi_pose_estimator_b200')
    synth = __import__('importlib').import_module('3d_multi_pose_estimator_b200.synth')
    cfg = b200.CameraConfig.from_parameters(parameters, name=config)
    cfg.to_npz(os.path.join(HERE, 'cameras_%s.npz' % config))
    n_feats = len(graph_generator.HumanGraphFromView.get_all_features('3'))
    gat, mlp = build_reference_models(parameters, n_feats)
    out = {}
    out['kinv32'] = np.stack([m.numpy() for m in graph_generator.inverse_camera_matrices])
    out['ti32'] = np.stack([m.numpy() for m in graph_generator.camera_i_transforms])
    out['centre32'] = np.stack([m.numpy() for m in graph_generator.all_cameras_from_root])
    cases = _cases(config)
    frames = {tag: synth.make_frame(cfg, seed, P, **kw) for tag, P, seed, kw in cases}
    calib_frames = [synth.make_frame(cfg, 1000 + i, 4 if config != 'arp3' else 8, all_cameras=config in ('arp_robot2', 'pansub')) for i in range(2)]
    gain, shift, z = calibrate(gat, calib_frames, parameters, mods)
    print(config, 'calibration gain %.4f shift %.6f -> logits std %.3f median %.4f' % (gain, shift, z.std(), np.median(z)))
    meta = dict(config=config, cases=[c[0] for c in cases], gat_seed=GAT_SEED, mlp_seed=MLP_SEED,
                mlp_in_dim=int(mlp.layers[1].in_features),
                calib_gain=gain, calib_shift=shift,
                gat_last_fc2_bias=float(gat.layers[-1].fc2.bias[0]),
                gat_checksum=checksum(gat.state_dict()), mlp_checksum=checksum(mlp.state_dict()),
                frames={tag: frames[tag] for tag in frames}, no_graph=[])
    out['gat_last_fc2_weight'] = gat.layers[-1].fc2.weight.numpy().copy()
    out['gat_last_fc2_bias'] = gat.layers[-1].fc2.bias.numpy().copy()
    graphs = []
    for tag, P, seed, kw in cases:
        rec = reference_frame(frames[tag], parameters, gat, mlp, mods)
        if rec is None:
            meta['no_graph'].append(tag)
            continue
        keep_layers = tag in (cases[0][0], 'rag1', 'rag')
        for k, v in rec.items():
            if k.startswith('gat_l') and not keep_layers:
                continue
            out['%s/%s' % (tag, k)] = v
        # graph again for the clustering fuzz
        pi = {c: [frames[tag][c][0], 0.0] for c in frames[tag] if json.loads(frames[tag][c][0])}
        sc = graph_generator.MergedMultipleHumansDataset(pi, mode='test', limit=10000, debug=True, alt='3', verbose=False)
        graphs.append((tag, sc.graphs[0], torch.squeeze(sc.data['edge_nodes_indices'][0]), sc.data['nodes_camera'][0]))
        print(config, tag, 'N=%d E=%d persons=%d' % (rec['n_nodes'], len(rec['src']), len(rec['proposals'])),
              'score range %.3f..%.3f' % (rec['scores'][rec['indices']].min(), rec['scores'][rec['indices']].max()))
    fuzz = cluster_fuzz(graphs, parameters, mods, np.random.default_rng(7), 240 if config in ('panoptic', 'arp3') else 60)
    meta['fuzz_tags'] = [f[0] for f in fuzz]
    for i, (tag, s, arr) in enumerate(fuzz):
        out['fuzz/%d/scores' % i] = s
        out['fuzz/%d/proposals' % i] = arr
    np.savez_compressed(os.path.join(HERE, 'golden_%s.npz' % config), **out)
    json.dump(meta, open(os.path.join(HERE, 'golden_%s.json' % config), 'w'))
    print(config, 'written', os.path.getsize(os.path.join(HERE, 'golden_%s.npz' % config)) // 1024, 'KiB')


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == '--one':
        run_config(sys.argv[2])
    else:
        for c in (sys.argv[1:] or ['panoptic', 'arp3', 'ring10']):
            subprocess.check_call([sys.executable, os.path.abspath(__file__), '--one', c])
