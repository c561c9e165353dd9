"""Attribution probe for the frame-resident aggregation kernel: one GAT layer-1-sized launch with parts switched off."""
import importlib, os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from bench import load_workload, load_weights
pm = importlib.import_module('3d_multi_pose_estimator_b200.pipeline')
pack = importlib.import_module('3d_multi_pose_estimator_b200.pack')
L = importlib.import_module('3d_multi_pose_estimator_b200._lib').lib()
# usage: agg_probe.py [frames config persons], e.g. 64 ring10 16 for the large-frame kernel
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
config = sys.argv[2] if len(sys.argv) > 2 else 'panoptic'
persons = int(sys.argv[3]) if len(sys.argv) > 3 else 4
cfg, frames = load_workload(config, n_frames, persons, 0)
gat, mlp = load_weights(config, cfg)
pb = pack.pack_frames(frames, cfg, keep_json=False)
pipe = pm.PosePipeline(cfg, gat, mlp, device='cuda:0')
pipe.agg_impl = int(sys.argv[4]) if len(sys.argv) > 4 else 0
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
# every 400-wide-or-narrower layer with planes output (the last, scalar layer has its own kernel)
for li, lay in enumerate(pipe.gat[:-1]):
    rows = db.n_heads + 1 if li == 0 else N
    z = torch.randn(rows, lay['ldz'], device='cuda') * 0.3
    act = pipe.planes_ws('probe_act%d' % li, N, lay['hd'])
    ts = []
    for i in range(7):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipe.aggregate(db, g, z, lay, layer0=(li == 0), raw=None, act=act, scores=None)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    print('layer %d (hd %d, ldz %d)  %.1f us' % (li, lay['hd'], lay['ldz'], 1e3 * sum(ts) / len(ts)))
# clustering phases of frame 0 (device printf under debug bit 128)
L.b200pose_set_debug(128)
try:
    out = pipe.infer(db)
    torch.cuda.synchronize()
finally:
    L.b200pose_set_debug(0)
