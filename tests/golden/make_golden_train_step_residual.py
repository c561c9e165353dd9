"""Goldens for the optimisation step with RESIDUAL GAT layers (gat2.py:43-48, 70-75; not the reference's shipped training
configuration, train_skeleton_matching.py:49): the unmodified reference GAT2(residual=True) under torch autograd, the loop body
of train_skeleton_matching.py:163-184, two steps on the first batch of tests/golden/make_golden_train_step.py for two models -
'resfc' (the shipped layer shapes: a res_fc projection on every layer after the first) and 'ident' (3 layers whose hidden layer
has in_dim == out_dim: the identity branch broadcast over 4 heads). Stored like golden_train_step.npz: losses, scores, a digest +
strided sample of every gradient, the parameters after the steps. oracle/train_oracle.py is checked against all of it first.

    python tests/golden/make_golden_train_step_residual.py        # writes golden_train_step_residual.npz
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
import make_golden as MG  # noqa: E402
import make_golden_training as MGT  # noqa: E402
import make_golden_train_step as MS  # noqa: E402

MODELS = {'resfc': ([40, 40, 40, 30], [10, 10, 8, 5], 17), 'ident': ([40, 40], [1, 4], 18)}      # hidden, heads, seed
STEPS = 2


def main():
    import torch
    parameters = MG._activate('panoptic')
    import gat2, graph_generator  # noqa: E401,F401
    import dgl
    import importlib
    b200 = importlib.import_module('3d_multi_pose_estimator_b200')
    synth = importlib.import_module('3d_multi_pose_estimator_b200.synth')
    weights_mod = importlib.import_module('3d_multi_pose_estimator_b200.weights')
    from oracle import train_oracle as TO
    cfg = b200.CameraConfig.from_parameters(parameters, name='panoptic')
    n_feats = len(graph_generator.HumanGraphFromView.get_all_features('3'))
    gm = json.load(open(os.path.join(HERE, 'golden_training_panoptic.json')))
    files = MGT.single_person_files(cfg, synth)
    paths = []
    for i, fr in enumerate(files):
        path = '/tmp/b200pose_single_%d.json' % i
        json.dump(fr, open(path, 'w'))
        paths.append(path)
    random.seed(gm['seed'])
    ds = graph_generator.MergedMultipleHumansDataset(paths, gm['probabilities'], limit=gm['limit'], mode='train', alt='3', raw_dir='.',
                                                     verbose=False, debug=True)
    members = list(range(MS.BATCH))
    subgraph, labels, indices = MS.collate([ds[i] for i in members], dgl, torch)
    feats = subgraph.ndata['h']
    src, dst = [t.numpy() for t in subgraph.edges()]
    out = {'members': np.array(members), 'indices': indices.numpy().ravel().astype(np.int64), 'labels': labels.numpy().ravel().astype(np.float32),
           'steps': np.array([STEPS])}
    for name, (hidden, heads, seed) in MODELS.items():
        torch.manual_seed(seed)
        model = gat2.GAT2(None, len(hidden) + 1, n_feats, 1, hidden, heads, torch.nn.LeakyReLU(), torch.nn.Sigmoid(), 0., 0., 0.15, True, bias=True)
        mine = weights_mod.make_gat_state(n_feats, seed, True, hidden, heads, residual=True)
        assert sorted(mine) == sorted(model.state_dict()) and all(torch.equal(mine[k], v) for k, v in model.state_dict().items())
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1.e-20)
        ow = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
        oadam = TO.Adam(ow)
        model.train()
        for step in range(STEPS):
            opt.zero_grad()
            model.g = subgraph
            for layer in model.layers:
                layer.g = subgraph
            outputs = torch.squeeze(model(feats.float(), subgraph))
            loss = torch.nn.MSELoss()(outputs[indices].float(), labels.float())
            loss.backward()
            pre = '%s/step%d/' % (name, step)
            out[pre + 'loss'] = np.array([loss.item()])
            out[pre + 'scores'] = outputs.detach().numpy().copy()
            oloss, oscores, ograds = TO.forward_backward(ow, feats.numpy(), src, dst, indices.numpy().ravel(), labels.numpy().ravel(),
                                                         heads=tuple(heads) + (1,), residual=True)
            assert abs(oloss - loss.item()) <= 1e-5 * loss.item()
            worst = 0.0
            for k, p in model.named_parameters():
                g = p.grad.numpy()
                d, s = MS.digest(g)
                out[pre + 'grad/' + k + '/digest'] = d
                out[pre + 'grad/' + k + '/sample'] = s
                err = np.abs(ograds[k].reshape(g.shape) - g).max()
                assert err <= 2e-4 * np.abs(g).max() or err <= 1e-8, (name, step, k, err)
                worst = max(worst, err / max(np.abs(g).max(), 1e-4))
            opt.step()
            oadam.step(ow, ograds)
            print(name, 'step', step, 'loss', loss.item(), 'worst oracle gradient error', worst)
        for k, v in model.state_dict().items():
            d, s = MS.digest(v.detach().numpy())
            out['%s/final/%s/digest' % (name, k)] = d
            out['%s/final/%s/sample' % (name, k)] = s
    np.savez_compressed(os.path.join(HERE, 'golden_train_step_residual.npz'), **out)
    print('written', os.path.getsize(os.path.join(HERE, 'golden_train_step_residual.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    main()
