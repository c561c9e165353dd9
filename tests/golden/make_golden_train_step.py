"""Goldens for the optimisation step of the skeleton-matching training loop (the part after the forward):
the UNMODIFIED reference (/root/reference + the import shims of oracle/shims) runs the loop body of
skeleton_matching/train_skeleton_matching.py:163-184 - zero_grad, GAT2 forward on a dgl.batch of training graphs, MSE loss on
the edge-node outputs, loss.backward() (torch autograd), torch.optim.Adam(lr=1e-4, weight_decay=1e-20).step() - for a few
steps, on the training dataset of tests/golden/make_golden_training.py.

    python tests/golden/make_golden_train_step.py        # writes golden_train_step.npz

Stored per step: loss, the scores of the batch, and of every gradient tensor its sum, absolute sum, largest magnitude and a
strided sample (small tensors whole); after the last step the same digest of every parameter. The restatement
oracle/train_oracle.py is checked against all of it here before anything is written.
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

GAT_SEED = 7
BATCH = 3          # graphs per step (the reference uses 15, train_skeleton_matching.py:42; the golden dataset holds ~10 graphs)
STEPS = 3
SAMPLE_STRIDE = 97


def digest(a):
    a = np.asarray(a, dtype=np.float32).ravel()
    sample = a if a.size <= 4096 else a[::SAMPLE_STRIDE]
    return np.array([a.astype(np.float64).sum(), np.abs(a.astype(np.float64)).sum(), np.abs(a).max()]), sample.copy()


def collate(batch, dgl, torch):
    """train_skeleton_matching.py:63-84"""
    graphs = [batch[0][0]]
    batched_labels = batch[0][1]
    batched_indices = batch[0][2]
    total_nodes = batch[0][0].number_of_nodes()
    for graph, labels, indices, _ in batch[1:]:
        graphs.append(graph)
        batched_labels = torch.cat([batched_labels, labels], dim=0)
        batched_indices = torch.cat([batched_indices, indices + total_nodes], dim=0)
        total_nodes += graph.number_of_nodes()
    return dgl.batch(graphs), batched_labels, batched_indices


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
    assert files == gm['files']
    paths = []
    for i, fr in enumerate(files):
        path = '/tmp/b200pose_single_%d.json' % i
        json.dump(fr, open(path, 'w'))
        paths.append(path)
    random.seed(gm['seed'])
    ds = graph_generator.MergedMultipleHumansDataset(paths, gm['probabilities'], limit=gm['limit'], mode='train', alt='3', raw_dir='.',
                                                     verbose=False, debug=True)
    # train_skeleton_matching.py:40-56, 148-152
    torch.manual_seed(GAT_SEED)
    model = gat2.GAT2(None, 5, n_feats, 1, [40, 40, 40, 30], [10, 10, 8, 5], torch.nn.LeakyReLU(), torch.nn.Sigmoid(), 0., 0., 0.15,
                      False, bias=True)
    state0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    mine = weights_mod.make_gat_state(n_feats, GAT_SEED, True)
    assert all(torch.equal(state0[k], mine[k]) for k in state0), 'seeded constructor order differs'
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1.e-20)
    loss_function = torch.nn.MSELoss()
    ow = {k: v.numpy().copy() for k, v in state0.items()}
    oadam = TO.Adam(ow)
    out = {}
    model.train()
    n_batches = len(ds) // BATCH
    assert n_batches >= 1
    for step in range(STEPS):
        b = step % n_batches
        subgraph, labels, indices = collate([ds[i] for i in range(b * BATCH, (b + 1) * BATCH)], dgl, torch)
        optimizer.zero_grad()
        feats = subgraph.ndata['h']
        model.g = subgraph
        for layer in model.layers:
            layer.g = subgraph
        outputs = torch.squeeze(model(feats.float(), subgraph))
        filtered_output = outputs[indices]
        loss = loss_function(filtered_output.float(), labels.float())
        loss.backward()
        pre = 'step%d/' % step
        out[pre + 'loss'] = np.array([loss.item()])
        out[pre + 'scores'] = outputs.detach().numpy().copy()
        out[pre + 'indices'] = indices.numpy().ravel().astype(np.int64)
        out[pre + 'labels'] = labels.numpy().ravel().astype(np.float32)
        out[pre + 'members'] = np.arange(b * BATCH, (b + 1) * BATCH)
        # ---- the restatement on the same batch, from its own copy of the parameters ----
        src, dst = [t.numpy() for t in subgraph.edges()]
        oloss, oscores, ograds = TO.forward_backward(ow, feats.numpy(), src, dst, indices.numpy().ravel(), labels.numpy().ravel())
        assert abs(oloss - loss.item()) <= 1e-5 * abs(loss.item()), (oloss, loss.item())
        worst = 0.0
        for k, p in model.named_parameters():
            g = p.grad.numpy()
            d, s = digest(g)
            out[pre + 'grad/' + k + '/digest'] = d
            out[pre + 'grad/' + k + '/sample'] = s
            aerr = np.abs(ograds[k].reshape(g.shape) - g).max()
            err = aerr / max(np.abs(g).max(), 1e-30)
            if aerr > 1e-9:        # (the last layer's attn_r gradient is a cancellation to ~1e-10: only the absolute error means anything)
                worst = max(worst, err)
            assert err <= 2e-4 or aerr <= 1e-9, ('oracle gradient', step, k, err, g.ravel()[:4], ograds[k].ravel()[:4])
        print('step', step, 'loss', loss.item(), 'N', subgraph.number_of_nodes(), 'M', len(indices), 'worst oracle gradient error (rel. to max)', worst)
        optimizer.step()
        oadam.step(ow, ograds)
        werr = max(np.abs(ow[k] - v.detach().numpy()).max() for k, v in model.state_dict().items())
        assert werr <= 2.5e-5, ('oracle parameters after step', step, werr)
        print('   parameters after the step: oracle vs reference max abs', werr)
    for k, v in model.state_dict().items():
        d, s = digest(v.detach().numpy())
        out['final/' + k + '/digest'] = d
        out['final/' + k + '/sample'] = s
    out['gat_seed'] = np.array([GAT_SEED])
    out['batch'] = np.array([BATCH])
    out['steps'] = np.array([STEPS])
    np.savez_compressed(os.path.join(HERE, 'golden_train_step.npz'), **out)
    print('written', os.path.getsize(os.path.join(HERE, 'golden_train_step.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    main()
