"""Debug probe: trainer forward / backward against the oracle, error per tensor (no asserts)."""
import importlib, json, os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import helpers
from oracle import train_oracle as TO
from oracle import pose_oracle as O
pm = importlib.import_module('3d_multi_pose_estimator_b200.pipeline')
pack_mod = importlib.import_module('3d_multi_pose_estimator_b200.pack')
tr = importlib.import_module('3d_multi_pose_estimator_b200.train')
cfg, npz, meta = helpers.load_golden('panoptic')
tags = helpers.graph_cases('panoptic')
frames = [{c: meta['frames'][t][c] for c in meta['frames'][t] if json.loads(meta['frames'][t][c][0])} for t in tags]
pb = pack_mod.pack_frames(frames, cfg)
db = pm.HostBatch(pb).to_device('cuda:0')
pipe = pm.PosePipeline(cfg, None, None, device='cuda:0')
g = pipe.build_graph(db, with_coo=True)
state = helpers.weights_mod.make_gat_state(cfg.n_features_sm, 11, True)
trainer = tr.GatTrainer(pipe, state)
node_off, head_off = pb.node_off, pb.head_off
idx = np.concatenate([np.arange(node_off[b] + (head_off[b + 1] - head_off[b]), node_off[b + 1]) for b in range(pb.n_frames)])
rng = np.random.default_rng(5)
labels = (rng.random(len(idx)) < 0.3).astype(np.float32)
d_idx = torch.from_numpy(idx.astype(np.int32)).cuda(); d_lab = torch.from_numpy(labels).cuda()
loss = float(trainer.step(db, g, d_idx, d_lab, update=False).item())
scores = trainer.last_scores.cpu().numpy().copy()
feats = pipe.node_features_f32(db).cpu().numpy()
row_ptr, col = g.row_ptr.cpu().numpy(), g.col[: db.n_edges].cpu().numpy()
src, dst = col, np.repeat(np.arange(db.n_nodes), np.diff(row_ptr))
w = helpers.np_state(state)
oloss, oscores, ograds = TO.forward_backward(w, feats, src, dst, idx, labels)
print('loss', loss, oloss)
print('scores rel err', (np.abs(scores - oscores) / np.abs(oscores)).max())
layers = pipe.prepare_gat(state)
x0 = pm.Planes.from_f32(pipe.node_features_f32(db), pipe._stream())
inf = pipe.gat_forward(db, g, x0=x0, dense_rows=True, layers=layers).cpu().numpy()
print('inference-path scores vs oracle', (np.abs(inf - oscores) / np.abs(oscores)).max(), ' trainer vs inference', np.abs(inf - scores).max())
# per-layer z against the oracle
h = feats
for l, c in enumerate(trainer.net.cache['layers']):
    p = TO.layer_params(w, l)
    out, cache = TO.layer_forward(h, src.astype(np.int64), dst.astype(np.int64), p, (10, 10, 8, 5, 1)[l], 0.15)
    z = c['z'][: db.n_nodes].cpu().numpy()
    hd = cache['ft2'].shape[1] * cache['ft2'].shape[2]
    H = cache['ft2'].shape[1]
    ft2 = cache['ft2'].reshape(db.n_nodes, -1)
    a1 = np.einsum('nhd,hd->nh', cache['ft2'], p['attn_l'][:, :, 0]); a2 = np.einsum('nhd,hd->nh', cache['ft2'], p['attn_r'][:, :, 0])
    print('layer', l, 'ft2 err', np.abs(z[:, :hd] - ft2).max() / np.abs(ft2).max(), 'a1 err', np.abs(z[:, hd:hd + H] - a1).max() / np.abs(a1).max(),
          'a2 err', np.abs(z[:, hd + H:hd + 2 * H] - a2).max() / np.abs(a2).max(),
          'h2 err', np.abs(c['h2'].to_f32()[: db.n_nodes].cpu().numpy() - cache['h2']).max() / np.abs(cache['h2']).max(),
          'x err', np.abs(c['x'].to_f32()[: db.n_nodes].cpu().numpy() - h).max())
    if l < 4:
        h = TO.leaky(out.reshape(db.n_nodes, -1), 0.01)
grads = {k: v.clone().cpu().numpy() for k, v in trainer.net.grads().items()}
for k, og in ograds.items():
    got = grads[k].reshape(og.shape)
    print('%-24s max|g| %.3e  err/max %.3e' % (k, np.abs(og).max(), np.abs(got - og).max() / max(np.abs(og).max(), 1e-30)))
# the same with the backward run by the oracle on the DEVICE's forward state (identical activation masks)
N = db.n_nodes
cl = trainer.net.cache['layers']
caches, raws = TO.caches_from_device(w, [c['x'].to_f32()[:N].cpu().numpy() for c in cl], [c['h2'].to_f32()[:N].cpu().numpy() for c in cl],
                                     [c['z'][:N].cpu().numpy() for c in cl], src, dst)
dl = trainer.net.buf.f('dlogit', N, 1)[:N, 0].cpu().numpy().reshape(-1, 1, 1)
g2 = TO.backward(w, caches, raws, src.astype(np.int64), dst.astype(np.int64), dl)
for k, og in ograds.items():
    got = grads[k].reshape(og.shape)
    dev = g2[k].reshape(og.shape)
    l2 = np.linalg.norm((got - og).ravel()) / max(np.linalg.norm(og.ravel()), 1e-30)
    nbad = int((np.abs(got - og) > 1e-4 * np.abs(og).max()).sum())
    print('%-22s vs oracle: max %.2e  L2 %.2e  n(>1e-4 max) %d of %d | vs oracle backward on device state: max %.2e' % (
        k, np.abs(got - og).max() / max(np.abs(og).max(), 1e-30), l2, nbad, og.size, np.abs(got - dev).max() / max(np.abs(dev).max(), 1e-30)))
