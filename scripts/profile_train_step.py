"""One profiled optimisation step of the training workload (for ncu): bench.py --workload train_step's batch (15 process_training
graphs by default), 3 warm steps, then cudaProfilerStart .. Stop around one launch-by-launch step (GatTrainer.step)."""
import importlib, os, random, sys
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, 'tests', 'golden')
pkg = importlib.import_module('3d_multi_pose_estimator_b200')
synth = importlib.import_module('3d_multi_pose_estimator_b200.synth')
pm = importlib.import_module('3d_multi_pose_estimator_b200.pipeline')
tg = importlib.import_module('3d_multi_pose_estimator_b200.training_graphs')
tr = importlib.import_module('3d_multi_pose_estimator_b200.train')
W = importlib.import_module('3d_multi_pose_estimator_b200.weights')
G = int(sys.argv[1]) if len(sys.argv) > 1 else 15
cfg = pkg.CameraConfig.from_npz(os.path.join(GOLDEN, 'cameras_panoptic.npz'))
files = [[synth.make_frame(cfg, 7000 + 100 * f + t, 1, drop_joint_p=0.1, drop_view_p=0.1) for t in range(24)] for f in range(4)]
random.seed(0)
inputs, indices = tg.load_inputs(files, 'train', cfg.used_pe_names, random)
built = []
for mp in tg.sample_sets(inputs, indices, [0.8, 0.6, 0.7, 0.5], 10 ** 9, random):
    b = tg.training_graph_inputs(mp, cfg)
    if b is not None:
        built.append(b)
    if len(built) >= min(G, 256):
        break
members = [built[i % len(built)] for i in range(G)]
pb, pairs = tg.batch_packed([(m[0], m[1]) for m in members])
idx, off = [], 0
for m in members:
    H, N = m[0].n_heads, int(m[0].node_off[-1])
    idx.append(np.arange(off + H, off + N)); off += N
idx = np.concatenate(idx).astype(np.int32)
labels = np.concatenate([m[2].ravel() for m in members]).astype(np.float32)
dev = torch.device('cuda:0')
pipe = pm.PosePipeline(cfg, None, None, device=dev)
trainer = tr.GatTrainer(pipe, W.make_gat_state(cfg.n_features_sm, 0, True))
db = pm.HostBatch(pb).to_device(dev)
g = pipe.build_graph_pairs(db, torch.from_numpy(pairs).to(dev), with_coo=False)
d_idx, d_lab = torch.from_numpy(idx).to(dev), torch.from_numpy(labels).to(dev)
x0 = trainer.features(db)
for _ in range(3):
    trainer.step(db, g, d_idx, d_lab, x0=x0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = trainer.step(db, g, d_idx, d_lab, x0=x0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('graphs', G, 'nodes', db.n_nodes, 'loss', float(loss.item()))
