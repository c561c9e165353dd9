"""Where the host time of GatTrainer.step_captured goes (wall clock per part, synchronised)."""
import importlib, os, sys, time, random
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import helpers
pkg = importlib.import_module('3d_multi_pose_estimator_b200')
pm = importlib.import_module('3d_multi_pose_estimator_b200.pipeline')
tg = importlib.import_module('3d_multi_pose_estimator_b200.training_graphs')
tr = importlib.import_module('3d_multi_pose_estimator_b200.train')
synth = helpers.synth
cfg, _, _ = helpers.load_golden('panoptic')
files = [[synth.make_frame(cfg, 7000 + 100 * f + t, 1, drop_joint_p=0.1, drop_view_p=0.1) for t in range(24)] for f in range(4)]
random.seed(0)
inputs, indices = tg.load_inputs(files, 'train', cfg.used_pe_names, random)
built = []
for mp in tg.sample_sets(inputs, indices, [0.8, 0.6, 0.7, 0.5], 10 ** 9, random):
    b = tg.training_graph_inputs(mp, cfg)
    if b is not None:
        built.append(b)
    if len(built) >= 15:
        break
pb, pairs = tg.batch_packed([(m[0], m[1]) for m in built])
idx, off = [], 0
for m in built:
    H, N = m[0].n_heads, int(m[0].node_off[-1]); idx.append(np.arange(off + H, off + N)); off += N
idx = np.concatenate(idx).astype(np.int32); labels = np.concatenate([m[2].ravel() for m in built]).astype(np.float32)
dev = torch.device('cuda:0')
pipe = pm.PosePipeline(cfg, None, None, device=dev)
trainer = tr.GatTrainer(pipe, helpers.weights_mod.make_gat_state(cfg.n_features_sm, 0, True))
hb = pm.HostBatch(pb); db = hb.to_device(dev)
g = pipe.build_graph_pairs(db, torch.from_numpy(pairs).to(dev), with_coo=False)
d_idx, d_lab = torch.from_numpy(idx).to(dev), torch.from_numpy(labels).to(dev)
x0 = trainer.features(db)
for _ in range(5):
    trainer.step_captured(db, g, d_idx, d_lab, x0=x0)
torch.cuda.synchronize()
def timed(name, fn, n=50):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print('%-40s host %.3f ms   with sync %.3f ms' % (name, 1e3 * (t1 - t0) / n, 1e3 * (t2 - t0) / n))
ent = list(trainer._graphs.values())[0]
timed('graph.replay()', lambda: ent['graph'].replay())
timed('step_captured(x0 given)', lambda: trainer.step_captured(db, g, d_idx, d_lab, x0=x0))
timed('step_captured(features inside)', lambda: trainer.step_captured(db, g, d_idx, d_lab))
timed('step (eager)', lambda: trainer.step(db, g, d_idx, d_lab, x0=x0))
timed('features', lambda: trainer.features(db))
timed('8 copies', lambda: [ent['x0'].hi.copy_(x0.hi, non_blocking=True) for _ in range(8)])
timed('hb.to_device + build_graph_pairs', lambda: pipe.build_graph_pairs(hb.to_device(dev), torch.from_numpy(pairs).to(dev), with_coo=False))
timed('step_captured + item', lambda: trainer.step_captured(db, g, d_idx, d_lab, x0=x0).item())
