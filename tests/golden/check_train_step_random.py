"""One-off pinning run (needs /root/reference; not part of the pytest suites): the training-step restatement
oracle/train_oracle.py against the UNMODIFIED reference GAT2 under torch autograd on 40 random cases - random layer counts,
widths and head counts (not only the shipped 10x40 / 10x40 / 8x40 / 5x30 / 1x1), with and without residual connections (res_fc and
identity branches), random symmetric graphs with self-loops and a few
hubs (in-degrees 3 .. 40), random features, labels on a random node subset - loss, scores and every gradient tensor of
forward + MSE + backward, then three Adam steps.

    python tests/golden/check_train_step_random.py      # prints the worst errors; exits non-zero above the tolerances
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402


def random_graph(rng, n):
    adj = np.eye(n, dtype=bool)
    for u in range(n):
        for v in rng.choice(n, 2, replace=False):
            adj[u, v] = adj[v, u] = True
    for h in rng.choice(n, 3, replace=False):
        for v in rng.choice(n, min(n, 40), replace=False):
            adj[h, v] = adj[v, h] = True
    dst, src = np.nonzero(adj)
    order = rng.permutation(len(src))                 # DGL keeps insertion order: any order must give the same gradients
    return src[order], dst[order]


def main():
    import torch
    MG._activate('panoptic')
    import dgl
    from gat2 import GAT2
    from oracle import train_oracle as TO
    rng = np.random.default_rng(21)
    worst_g, worst_s, worst_l, worst_w, worst_at = 0.0, 0.0, 0.0, 0.0, None
    n_identity = 0
    for case in range(40):
        n_hidden = int(rng.integers(1, 4))
        hidden = [int(rng.integers(2, 24)) for _ in range(n_hidden)]
        heads = [int(rng.integers(1, 6)) for _ in range(n_hidden)]
        residual = case % 2 == 1                          # every other model with residual connections (gat2.py:43-48, 70-75) ...
        if residual and n_hidden >= 2 and case % 4 == 1:  # ... half of those with an identity branch: in_dim == out_dim of layer 1
            heads[0] = 1
            hidden[1] = hidden[0]
        in_dim = int(rng.integers(5, 60))
        n = int(rng.integers(45, 140))
        src, dst = random_graph(rng, n)
        feats = rng.standard_normal((n, in_dim)).astype(np.float32)
        idx = np.sort(rng.choice(n, int(rng.integers(5, n // 2)), replace=False))
        labels = (rng.random(len(idx)) < 0.4).astype(np.float32)
        torch.manual_seed(1000 + case)
        model = GAT2(None, n_hidden + 1, in_dim, 1, hidden, heads, torch.nn.LeakyReLU(), torch.nn.Sigmoid(), 0., 0., 0.15, residual, bias=True)
        n_identity += int(residual and any(getattr(l, 'res_fc', 1) is None for l in model.layers))
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1.e-20)
        ow = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
        oadam = TO.Adam(ow)
        g = dgl.graph((torch.from_numpy(src), torch.from_numpy(dst)), num_nodes=n, idtype=torch.int32)
        model.train()
        for step in range(3):
            opt.zero_grad()
            model.g = g
            for layer in model.layers:
                layer.g = g
            out = torch.squeeze(model(torch.from_numpy(feats), g))
            loss = torch.nn.MSELoss()(out[torch.from_numpy(idx)].float(), torch.from_numpy(labels))
            loss.backward()
            oloss, oscores, ograds = TO.forward_backward(ow, feats, src, dst, idx, labels, heads=tuple(heads) + (1,), residual=residual)
            worst_l = max(worst_l, abs(oloss - loss.item()) / loss.item())
            worst_s = max(worst_s, float(np.abs(oscores - out.detach().numpy()).max()))
            for k, p in model.named_parameters():
                gt = p.grad.numpy()
                err = np.abs(ograds[k].reshape(gt.shape) - gt).max()
                # relative to the tensor's maximum, floored by the scale of its terms: the last layer's attn_r gradient is a sum
                # that cancels to ~1e-10 (a2[v] shifts every logit of destination v alike) while its terms are ~1e-4
                if k == 'layers.%d.attn_r' % n_hidden:
                    assert err <= 1e-7, (case, step, k, err)
                    continue
                rel = err / max(np.abs(gt).max(), 1e-4)
                if rel > worst_g:
                    worst_g, worst_at = rel, (case, step, k, float(np.abs(gt).max()), float(err))
            opt.step()
            oadam.step(ow, ograds)
            worst_w = max(worst_w, max(float(np.abs(ow[k] - v.detach().numpy()).max()) for k, v in model.state_dict().items()))
    print('40 random models (20 with residual layers, %d of them with an identity branch) x 3 steps:' % n_identity, ' loss rel %.2e, scores abs %.2e, gradients (of each tensor maximum) %.2e, parameters abs %.2e'
          % (worst_l, worst_s, worst_g, worst_w), '; worst gradient at', worst_at)
    ok = worst_l <= 1e-5 and worst_s <= 1e-5 and worst_g <= 1e-3 and worst_w <= 2e-4
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
