"""CPU: the training-step restatement (oracle/train_oracle.py: forward, hand-written backward, Adam) against the goldens the
unmodified reference produced under torch autograd (tests/golden/make_golden_train_step.py: the loop body of
skeleton_matching/train_skeleton_matching.py:163-184 for three steps)."""
import json
import os
import random

import numpy as np

import helpers
from oracle import pose_oracle as O
from oracle import train_oracle as TO

GOLDEN = helpers.GOLDEN
STRIDE = 97


def sample(a):
    a = np.asarray(a, dtype=np.float32).ravel()
    return a if a.size <= 4096 else a[::STRIDE]


def golden_batches():
    """The batched training graphs of the golden steps, rebuilt by the oracle's dataset restatement from the same files/seed."""
    cfg, _, _ = helpers.load_golden('panoptic')
    gm = json.load(open(os.path.join(GOLDEN, 'golden_training_panoptic.json')))
    gz = np.load(os.path.join(GOLDEN, 'golden_train_step.npz'))
    tabs = O.CameraTables(cfg)
    random.seed(gm['seed'])
    inputs, indices = O.load_training_inputs([json.loads(json.dumps(f)) for f in gm['files']], 'train', cfg.used_pe_names, random)
    graphs = [g for g in (O.build_training_graph(mp, tabs) for mp in O.training_samples(inputs, indices, gm['probabilities'], gm['limit'], random))
              if g is not None]
    out = []
    for step in range(int(gz['steps'][0])):
        members = [graphs[i] for i in gz['step%d/members' % step]]
        bg = O.batch_graphs(members)
        idx, off = [], 0
        for m in members:
            idx.append(np.asarray(m['indices']) + off)
            off += m['n_nodes']
        labels = np.concatenate([m['labels'].ravel() for m in members]).astype(np.float32)
        out.append((bg, np.concatenate(idx), labels))
    return cfg, gz, out


def check_digest(name, got, gz, key, rtol, l2tol=None, atol=1e-9):
    """A tensor against its stored digest [sum, abs-sum, max-abs] and strided sample: largest sampled error relative to the
    tensor's maximum (absolute floor `atol`: the last layer's attn_r gradient cancels to ~1e-12 while its terms are ~1e-4), optionally the relative L2 error of the sample."""
    got = np.asarray(got, dtype=np.float32)
    dg, smp = gz[key + '/digest'], gz[key + '/sample']
    scale = max(float(dg[2]), 1e-30)
    diff = sample(got) - smp
    err = np.abs(diff).max()
    assert err <= rtol * scale or err <= atol, (name, err, scale)
    if l2tol is not None:
        l2 = np.linalg.norm(diff.astype(np.float64))
        assert l2 <= l2tol * np.linalg.norm(smp.astype(np.float64)) or err <= atol, (name, 'L2', l2)
    assert abs(np.abs(got.astype(np.float64)).sum() - dg[1]) <= max(1e-3, rtol) * dg[1] + (1e-7 * scale + atol) * got.size, (name, 'abs-sum')
    return 0.0 if err <= atol else err / scale


def check_parameters(state, gz, lr, steps, tight=2.5e-5, frac=0.05, prefix='final/'):
    """Parameters after `steps` Adam steps. Adam's first updates are lr * g / (|g| + eps): an element whose gradient is below
    the noise floor of the arithmetic moves by up to lr per step in EITHER direction, so two correct implementations that differ
    in the last bits of the forward agree tightly on most elements and by at most 2 * lr * steps on the rest. Returns the fraction
    of sampled elements further than `tight` from the reference."""
    n_far, n_all, worst = 0, 0, 0.0
    for k, v in state.items():
        d = np.abs(sample(np.asarray(v)) - gz[prefix + k + '/sample'])
        worst = max(worst, float(d.max()))
        n_far += int((d > tight).sum())
        n_all += d.size
    assert worst <= 2 * lr * steps + 1e-6, ('parameter after the steps', worst)
    assert n_far <= frac * n_all, ('parameters further than %g from the reference: %d of %d' % (tight, n_far, n_all))
    return n_far / n_all, worst


def test_training_step_restatement_against_reference_autograd():
    cfg, gz, batches = golden_batches()
    w = helpers.np_state(helpers.weights_mod.make_gat_state(cfg.n_features_sm, int(gz['gat_seed'][0]), True))
    adam = TO.Adam(w)
    worst = 0.0
    for step, (bg, idx, labels) in enumerate(batches):
        pre = 'step%d/' % step
        assert np.array_equal(idx, gz[pre + 'indices']) and np.array_equal(labels, gz[pre + 'labels'])
        loss, scores, grads = TO.forward_backward(w, bg['feats'], bg['src'], bg['dst'], idx, labels)
        ref_loss = float(gz[pre + 'loss'][0])
        assert abs(loss - ref_loss) <= 1e-5 * ref_loss, (step, loss, ref_loss)
        assert (np.abs(scores - gz[pre + 'scores']) / np.abs(gz[pre + 'scores'])).max() <= 1e-4
        for k, g in grads.items():
            worst = max(worst, check_digest('grad %s step %d' % (k, step), g, gz, pre + 'grad/' + k, 2e-4))
        adam.step(w, grads)
    for k, v in w.items():
        err = np.abs(sample(v) - gz['final/' + k + '/sample']).max()
        assert err <= 2.5e-5, ('parameter after %d steps' % len(batches), k, err)       # lr = 1e-4: a quarter of one step
    print('worst gradient error relative to the tensor maximum:', worst)


RESIDUAL_MODELS = {'resfc': ([40, 40, 40, 30], [10, 10, 8, 5], 17), 'ident': ([40, 40], [1, 4], 18)}     # make_golden_train_step_residual.py


def residual_batch():
    cfg, gz, batches = golden_batches()
    rz = np.load(os.path.join(GOLDEN, 'golden_train_step_residual.npz'))
    bg, idx, labels = batches[0]
    assert np.array_equal(idx, rz['indices']) and np.array_equal(labels, rz['labels'])
    return cfg, rz, bg, idx, labels


def test_residual_training_step_restatement_against_reference_autograd():
    """The same for GAT2(residual=True): res_fc layers and the identity branch (gat2.py:43-48, 70-75)."""
    cfg, rz, bg, idx, labels = residual_batch()
    for name, (hidden, heads, seed) in RESIDUAL_MODELS.items():
        w = helpers.np_state(helpers.weights_mod.make_gat_state(cfg.n_features_sm, seed, True, hidden, heads, residual=True))
        adam = TO.Adam(w)
        for step in range(int(rz['steps'][0])):
            pre = '%s/step%d/' % (name, step)
            loss, scores, grads = TO.forward_backward(w, bg['feats'], bg['src'], bg['dst'], idx, labels, heads=tuple(heads) + (1,), residual=True)
            ref_loss = float(rz[pre + 'loss'][0])
            assert abs(loss - ref_loss) <= 1e-5 * ref_loss, (name, step, loss, ref_loss)
            assert sorted(grads) == sorted(w)
            for k, g in grads.items():
                check_digest('grad %s %s step %d' % (name, k, step), g, rz, pre + 'grad/' + k, 2e-4, atol=1e-8)
            adam.step(w, grads)
        # (the last layer's attn_r has a gradient that cancels to ~1e-10: Adam's g / |g| moves it by lr per step in a direction the
        # last bit decides - check_parameters bounds that by 2 lr steps and asks for tight agreement on all but a sliver)
        check_parameters(w, rz, 1e-4, int(rz['steps'][0]), frac=0.001, prefix=name + '/final/')
