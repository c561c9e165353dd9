"""Attribution probe for the frame-resident aggregation kernel: one GAT layer-1-sized launch with parts switched off."""
import importlib, os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from bench import load_workload, load_weights
pm = importlib.import_module('3d_multi_pose_estimator_b200.pipeline')
pack = importlib.import_module('3d_multi_pose_estimator_b200.pack')
L = importlib.import_module('3d_multi_pose_estimator_b200._lib').lib()
cfg, frames = load_workload('panoptic', 1024, 4, 0)
gat, mlp = load_weights('panoptic', cfg)
pb = pack.pack_frames(frames, cfg, keep_json=False)
pipe = pm.PosePipeline(cfg, gat, mlp, device='cuda:0')
db = pm.HostBatch(pb).to_device('cuda:0')
g = pipe.build_graph(db, with_coo=False)
lay = pipe.gat[1]
N = db.n_nodes
z = torch.randn(N, lay['ldz'], device='cuda') * 0.3
act = pipe.planes_ws('probe_act', N, lay['hd'])
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
for dbg, name in ((0, 'full'), (16, 'no stores'), (32, 'no head contributions'), (64, 'no edge-node destinations'), (96, 'phases 1 + ring only'), (112, 'phases 1 + ring, no stores')):
    L.b200pose_set_debug(dbg)
    ts = []
    for i in range(7):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipe.aggregate(db, g, z, lay, layer0=False, raw=None, act=act, scores=None)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    print('%-32s %.1f us' % (name, 1e3 * sum(ts) / len(ts)))
L.b200pose_set_debug(0)
