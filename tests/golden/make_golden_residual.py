"""Regenerates tests/golden/golden_residual.npz: the UNMODIFIED reference GAT2 (skeleton_matching/gat2.py, CPU, dgl shim)
built with residual=True, run on graphs of the Panoptic goldens.

    python tests/golden/make_golden_residual.py

Two models: 'resfc' = the shipped layer shapes with residual=True (every layer after the first has in_dim != out_dim, so
each gets a res_fc projection, gat2.py:43-46, 70-72); 'ident' = a 3-layer model whose hidden layer has in_dim == out_dim
(heads[0] = 1), the identity branch that broadcasts the layer input over the attention heads (gat2.py:47-48, 73-74).
Weights come from the reference constructors under torch.manual_seed, so the GPU box rebuilds them with the drop-in's
constructors (same initialisation order); only scores and per-layer outputs are stored. Build container only.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

MODELS = {
    # name: (num_layers, num_hidden, heads, seed)
    'resfc': (5, [40, 40, 40, 30], [10, 10, 8, 5], 3),
    'ident': (3, [40, 40], [1, 4], 4),
}
TAGS = ['p3', 'rag1']


def main():
    from oracle import ref_env
    ref_env.activate_config('panoptic')
    import torch
    import dgl
    from gat2 import GAT2
    torch.set_grad_enabled(False)
    npz = np.load(os.path.join(HERE, 'golden_panoptic.npz'))
    out = {}
    for name, (n_layers, hidden, heads, seed) in MODELS.items():
        n_feats = npz[TAGS[0] + '/feats'].shape[1]
        torch.manual_seed(seed)
        model = GAT2(None, n_layers, n_feats, 1, hidden, heads, torch.nn.LeakyReLU(), torch.nn.Sigmoid(), 0., 0., 0.15, True,
                     bias=True)
        model.eval()
        for tag in TAGS:
            feats = torch.from_numpy(npz[tag + '/feats'])
            g = dgl.graph((torch.from_numpy(npz[tag + '/src']), torch.from_numpy(npz[tag + '/dst'])),
                          num_nodes=feats.shape[0], idtype=torch.int32)
            layers = []
            hooks = [lyr.register_forward_hook(lambda m, i, o, acc=layers: acc.append(o.detach().numpy().copy()))
                     for lyr in model.layers]
            scores = torch.squeeze(model(feats, g)).numpy()
            for h in hooks:
                h.remove()
            out['%s/%s/scores' % (name, tag)] = scores.astype(np.float32)
            for l, a in enumerate(layers):
                out['%s/%s/layer%d' % (name, tag, l)] = a.reshape(a.shape[0], -1).astype(np.float32)
        out[name + '/checksum'] = np.array([float(sum(p.double().sum() for p in model.parameters()))])
        out[name + '/keys'] = np.array(sorted(model.state_dict().keys()))
    np.savez_compressed(os.path.join(HERE, 'golden_residual.npz'), **out)
    print('wrote golden_residual.npz:', {k: v.shape for k, v in out.items() if k.endswith('scores')})


if __name__ == '__main__':
    main()
